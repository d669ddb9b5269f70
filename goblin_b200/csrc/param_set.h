// ParamSet: typed key/value bags filled from one JSON object, with the
// reference's lookup semantics (src/GoblinParamSet.cpp, parseParamSet in
// src/GoblinContextLoader.cpp:33-65):
//   * bool / integer / floating / string / 2-, 3-, 4-element arrays each go to
//     their own bag; other values are ignored;
//   * getX looks only in the bag of type X: a JSON integer is invisible to
//     getFloat (so "radius": 2 yields the default radius), and vice versa;
//   * the first entry with a matching key wins.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "host_math.h"
#include "json_reader.h"

namespace gb {

class ParamSet {
public:
    ParamSet() = default;
    explicit ParamSet(const JsonValue& object) { parse(object); }

    void parse(const JsonValue& object) {
        if (!object.isObject()) return;
        for (const auto& kv : object.obj) { // sorted-key order, like nlohmann
            const std::string& key = kv.first;
            const JsonValue& v = kv.second;
            switch (v.type) {
            case JsonValue::Bool: mBools.emplace_back(key, v.b); break;
            case JsonValue::Int: mInts.emplace_back(key, static_cast<int>(v.i)); break;
            case JsonValue::Float: mFloats.emplace_back(key, static_cast<float>(v.d)); break;
            case JsonValue::String: mStrings.emplace_back(key, v.s); break;
            case JsonValue::Array: {
                bool numeric = true;
                for (const JsonValue& e : v.arr) numeric = numeric && e.isNumber();
                if (!numeric) break;
                if (v.arr.size() == 2) {
                    Vec2 a; a.x = v.arr[0].asFloat(); a.y = v.arr[1].asFloat();
                    mVec2s.emplace_back(key, a);
                } else if (v.arr.size() == 3) {
                    mVec3s.emplace_back(key, Vec3(v.arr[0].asFloat(), v.arr[1].asFloat(), v.arr[2].asFloat()));
                } else if (v.arr.size() == 4) {
                    Vec4 a; a.x = v.arr[0].asFloat(); a.y = v.arr[1].asFloat();
                    a.z = v.arr[2].asFloat(); a.w = v.arr[3].asFloat();
                    mVec4s.emplace_back(key, a);
                }
                break;
            }
            default: break;
            }
        }
    }

    void setString(const std::string& k, const std::string& v) { mStrings.emplace_back(k, v); }
    void setFloat(const std::string& k, float v) { mFloats.emplace_back(k, v); }
    void setBool(const std::string& k, bool v) { mBools.emplace_back(k, v); }

    bool hasString(const std::string& k) const { return has(mStrings, k); }
    bool hasVector3(const std::string& k) const { return has(mVec3s, k); }
    bool hasInt(const std::string& k) const { return has(mInts, k); }
    bool hasFloat(const std::string& k) const { return has(mFloats, k); }

    bool getBool(const std::string& k, bool d = false) const { return get(mBools, k, d); }
    int getInt(const std::string& k, int d = 0) const { return get(mInts, k, d); }
    float getFloat(const std::string& k, float d = 0.0f) const { return get(mFloats, k, d); }
    Vec2 getVector2(const std::string& k, const Vec2& d = Vec2()) const { return get(mVec2s, k, d); }
    Vec3 getVector3(const std::string& k, const Vec3& d = Vec3()) const { return get(mVec3s, k, d); }
    Vec4 getVector4(const std::string& k, const Vec4& d = Vec4()) const { return get(mVec4s, k, d); }
    std::string getString(const std::string& k, const std::string& d = "") const { return get(mStrings, k, d); }

private:
    template <typename T>
    static bool has(const std::vector<std::pair<std::string, T>>& bag, const std::string& k) {
        for (const auto& e : bag) if (e.first == k) return true;
        return false;
    }
    template <typename T>
    static T get(const std::vector<std::pair<std::string, T>>& bag, const std::string& k, const T& d) {
        for (const auto& e : bag) if (e.first == k) return e.second;
        return d;
    }
    std::vector<std::pair<std::string, bool>> mBools;
    std::vector<std::pair<std::string, int>> mInts;
    std::vector<std::pair<std::string, float>> mFloats;
    std::vector<std::pair<std::string, std::string>> mStrings;
    std::vector<std::pair<std::string, Vec2>> mVec2s;
    std::vector<std::pair<std::string, Vec3>> mVec3s;
    std::vector<std::pair<std::string, Vec4>> mVec4s;
};

// getQuaternion / getTransform, src/GoblinUtils.cpp:71-91
inline Quat getQuaternion(const ParamSet& p) {
    if (p.hasVector3("euler")) {
        return eulerToQuat(p.getVector3("euler"), p.getString("rotation_order", "xyz"));
    }
    Vec4 d; d.x = 1; d.y = 0; d.z = 0; d.w = 0;
    Vec4 q = p.getVector4("orientation", d);
    return Quat(q.x, q.y, q.z, q.w);
}

inline Transform getTransform(const ParamSet& p) {
    Transform t;
    t.position = p.getVector3("position", Vec3(0, 0, 0));
    t.orientation = getQuaternion(p);
    t.scale = p.getVector3("scale", Vec3(1, 1, 1));
    t.update();
    return t;
}

} // namespace gb
