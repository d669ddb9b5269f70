// Radiance maps for image based lights: the OpenEXR reader behind the reference's
// loadImage (src/GoblinImageIO.cpp:14-34,128-144; the reference delegates to the
// vendored tinyexr's LoadEXR), the MIPMap pyramid (src/GoblinTexture.cpp:39-71)
// with its gaussian resizeImage (:520-597), and the CDF1D / CDF2D tables
// (src/GoblinSampler.cpp:309-394) an ImageBasedLight samples from.  Host code,
// same operation order and libm as the reference, so the tables are the
// reference's bit for bit.
#pragma once
#include <string>
#include <vector>

namespace gb {

// RGBA float pixels, top row first, as LoadEXR returns them: R, G, B by channel name, A = 1 when
// the file has none, a single-channel file broadcast to all four.  Scanline files, one part,
// compression none / RLE / ZIPS / ZIP, HALF / FLOAT / UINT channels.
bool loadEXR(const std::string& path, int* width, int* height, std::vector<float>* rgba, std::string* error);

// resizeImage<Color>: separable truncated gaussian, clamp addressing; alpha of the result is 1.
void resizeImage(const float* src, int srcWidth, int srcHeight, int dstWidth, int dstHeight, std::vector<float>* dst);

struct MipLevel { int width = 0, height = 0; std::vector<float> rgba; };
// MIPMap<Color>::MIPMap: level 0 is the image resized up to powers of two when needed; each
// further level is resizeImage of the previous one down to 1 x 1.
void buildMipmap(std::vector<float> rgba, int width, int height, std::vector<MipLevel>* levels);
// MIPMap::lookup(level, s, t) with repeat addressing (src/GoblinTexture.cpp:274-288)
void mipLookup(const std::vector<MipLevel>& levels, int level, float s, float t, float out[4]);

// CDF2D over a width x height function, flattened for gb_scene_desc.light_dist:
//   rows:     height x width function values, then height x (width + 1) normalised CDFs
//   marginal: height row integrals (its function), height + 1 CDF entries, 1 integral
void buildDistribution2D(const float* f2D, int width, int height, std::vector<float>* out);

} // namespace gb
