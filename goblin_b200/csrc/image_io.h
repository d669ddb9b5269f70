// Image output for the film: the reference writes colour / weight as an
// OpenEXR file with HALF B, G, R channels (src/GoblinImageIO.cpp:35-98) or a
// gamma-2.2 ASCII PPM (:100-126).  Here: an uncompressed scanline EXR with the
// same channel layout, the same PPM, and a raw float PFM for parity checks
// (HALF quantisation would pollute a z-test).
#pragma once
#include <string>

namespace gb {

// rgbw: yres x xres x 4 floats (weighted r, g, b, weight); the written colour
// is rgb / weight (Film::writeImage, src/GoblinFilm.cpp:164-173).
bool writeFilm(const std::string& path, const float* rgbw, int xres, int yres, std::string* error);

// rgb: yres x xres x 3 floats, already colour / weight (and bloomed).  toneMapping applies
// Goblin::toneMapping (src/GoblinImageIO.cpp:220-237) before a .ppm is written, and only then
// (Goblin::writeImage, :146-167).
bool writeRgb(const std::string& path, const float* rgb, int xres, int yres, bool toneMapping, std::string* error);

// The (2 * filterWidth - 1)^2 bloom kernel's quadrant table, filterWidth^2 entries
// (src/GoblinImageIO.cpp:174-182); filterWidth = ceil(bloomRadius * max(w, h)) / 2.
int bloomFilterWidth(float bloomRadius, int xres, int yres);
void bloomFilterTable(int filterWidth, float* table);

} // namespace gb
