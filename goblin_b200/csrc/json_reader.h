// Minimal JSON reader for Goblin scene files.
//
// The reference vendors nlohmann::json (src/json.hpp) purely as an I/O
// dependency.  What the loader relies on, and what is kept here:
//   * integer vs floating literals are distinct types (a literal is floating
//     iff it contains '.', 'e' or 'E'), because ParamSet never converts
//     between them (src/GoblinContextLoader.cpp:41-46);
//   * floating literals are parsed as double (strtod) and narrowed to float;
//   * objects iterate in sorted-key order; a repeated key keeps the last value.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace gb {

class JsonValue {
public:
    enum Type { Null, Bool, Int, Float, String, Array, Object };
    Type type = Null;
    bool b = false;
    int64_t i = 0;
    double d = 0.0;
    std::string s;
    std::vector<JsonValue> arr;
    std::map<std::string, JsonValue> obj;

    bool isObject() const { return type == Object; }
    bool isArray() const { return type == Array; }
    bool isNumber() const { return type == Int || type == Float; }
    float asFloat() const { return type == Int ? static_cast<float>(i) : static_cast<float>(d); }
    const JsonValue* find(const std::string& key) const {
        if (type != Object) return nullptr;
        auto it = obj.find(key);
        return it == obj.end() ? nullptr : &it->second;
    }
};

// Returns false and fills *error on malformed input.
bool parseJson(const std::string& text, JsonValue* out, std::string* error);

} // namespace gb
