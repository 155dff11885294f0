// stand-in for <yaml-cpp/yaml.h>: include/read_configs.h defines the config PODs (CFConfig, LoopClosureConfig, ...) next to an
// inline YAML loader; oracle/_ref fills the PODs directly, so the loader only has to compile.
#pragma once
#include <string>
namespace YAML {
class Node {
 public:
  Node operator[](const char*) const { return Node(); }
  Node operator[](const std::string&) const { return Node(); }
  template <class T> T as() const { return T(); }
};
inline Node LoadFile(const std::string&) { return Node(); }
}  // namespace YAML
