// TEST INFRASTRUCTURE.  Member definitions of the reference's GlobalConfig (declared in its own
// include/global_config.hpp, which src/motion_planning.cpp and the drop-in include) without yaml-cpp: load_file
// goes through the YAML-subset reader of toy-example-of-ilqr_b200/host/scenario_host.hpp, which fills the same
// "section/key" map with the same types and defaults as src/global_config.cpp:22-92.
#include <iostream>
#include <string>

#include "global_config.hpp"
#include "scenario_host.hpp"

GlobalConfig* GlobalConfig::instance = nullptr;

void GlobalConfig::load_file(const std::string& path) {
    cilqr_host::GlobalConfig reader;
    reader.load_file(path);
    config_map = reader.entries();
}
bool GlobalConfig::has_key(std::string key_str) { return config_map.find(key_str) != config_map.end(); }
GlobalConfig* GlobalConfig::get_instance(const std::string& path) {
    if (instance == nullptr) {
        instance = new GlobalConfig();
        instance->load_file(path);
    }
    return instance;
}
template <typename T>
T GlobalConfig::get_config(const std::string& key) const {
    auto it = config_map.find(key);
    if (it != config_map.end()) {
        try {
            return std::any_cast<T>(it->second);
        } catch (const std::bad_any_cast&) {
            std::cerr << "Type mismatch for key: " << key << std::endl;
        }
    } else {
        std::cerr << "Configuration key not found: " << key << std::endl;
    }
    return T();
}
void GlobalConfig::destroy_instance() {
    delete instance;
    instance = nullptr;
}
template std::string GlobalConfig::get_config<std::string>(const std::string&) const;
template int GlobalConfig::get_config<int>(const std::string&) const;
template double GlobalConfig::get_config<double>(const std::string&) const;
template bool GlobalConfig::get_config<bool>(const std::string&) const;
template std::vector<double> GlobalConfig::get_config<std::vector<double>>(const std::string&) const;
template std::vector<std::vector<double>> GlobalConfig::get_config<std::vector<std::vector<double>>>(const std::string&) const;
