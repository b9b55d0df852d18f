// STUB of boost::algorithm::trim (Boost 1.6x) for compiling the reference in place.
#pragma once
#include <string>
namespace boost { namespace algorithm {
inline void trim(std::string &s) {
    const char *ws = " \t\r\n\v\f";
    size_t b = s.find_first_not_of(ws), e = s.find_last_not_of(ws);
    s = b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
}
inline std::string trim_copy(std::string s) { trim(s); return s; }
} using algorithm::trim; using algorithm::trim_copy; }
