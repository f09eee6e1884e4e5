// boost/format.hpp STAND-IN (test infrastructure): the reference only formats log lines for its viewer with it.
#pragma once
#include <ostream>
#include <sstream>
#include <string>
namespace boost {
class format {
public:
  explicit format(const std::string& f) : text(f) {}
  template <class T>
  format& operator%(const T& v) {
    std::ostringstream os;
    os << " [" << v << "]";
    text += os.str();
    return *this;
  }
  friend std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.text; }

private:
  std::string text;
};
}  // namespace boost
