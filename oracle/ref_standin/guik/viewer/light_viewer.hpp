// guik/viewer/light_viewer.hpp STAND-IN (test infrastructure): the reference logs its progress into the Iridescence
// viewer; here the text is dropped.
#pragma once
#include <string>
namespace guik {
class LightViewer {
public:
  static LightViewer* instance() {
    static LightViewer v;
    return &v;
  }
  void append_text(const std::string&) {}
};
}  // namespace guik
