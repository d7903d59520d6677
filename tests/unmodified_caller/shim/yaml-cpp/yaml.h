// TEST INFRASTRUCTURE.  src/motion_planning.cpp includes <yaml-cpp/yaml.h> but uses nothing from it (the YAML is
// read by GlobalConfig); yaml-cpp is absent from this image, so the include resolves to this empty header.
#pragma once
