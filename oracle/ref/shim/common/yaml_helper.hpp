// Shadows legkilo/src/common/yaml_helper.hpp (a thin wrapper over yaml-cpp, which is not in this image): the same
// YamlHelper::get<T>(key[, default]) surface over a key -> numbers registry that the harness fills before it constructs
// KILO (KILO.cc:25-84 reads every option through this class).
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
namespace legkilo {
inline std::map<std::string, std::vector<double>>& yaml_registry() {
    static std::map<std::string, std::vector<double>> r;
    return r;
}
template <class T>
struct is_std_vector : std::false_type {};
template <class T>
struct is_std_vector<std::vector<T>> : std::true_type {};

class YamlHelper {
   public:
    YamlHelper() = delete;
    explicit YamlHelper(const std::string&) {}
    bool hasKey(const std::string& key) const { return yaml_registry().count(key) != 0; }
    template <typename T>
    T get(const std::string& key) const {
        auto it = yaml_registry().find(key);
        if (it == yaml_registry().end()) throw std::runtime_error("Failed to find key: " + key);
        return convert<T>(it->second);
    }
    template <typename T>
    T get(const std::string& key, const T& default_value) const {
        auto it = yaml_registry().find(key);
        if (it == yaml_registry().end()) return default_value;
        return convert<T>(it->second);
    }
   private:
    template <typename T>
    static T convert(const std::vector<double>& v) {
        if constexpr (is_std_vector<T>::value) {
            T out;
            for (double d : v) out.push_back(static_cast<typename T::value_type>(d));
            return out;
        } else if constexpr (std::is_same<T, bool>::value) {
            return v.at(0) != 0.0;
        } else {
            return static_cast<T>(v.at(0));
        }
    }
};
}  // namespace legkilo
