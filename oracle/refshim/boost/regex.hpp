// Stand-in for <boost/regex.hpp> (see ../README.md): std::regex (ECMAScript) under the boost name.  Boost's
// Perl dialect reads "[^]]" as "anything but ']'"; ECMAScript would close the class at the first ']', so the
// pattern text is rewritten before it reaches std::regex.
#pragma once
#include <regex>
#include <string>
namespace boost {
inline std::string refshim_translate(std::string p) {
    for (std::string::size_type i = p.find("[^]"); i != std::string::npos; i = p.find("[^]", i + 4)) p.replace(i, 3, "[^\\]");
    return p;
}
class regex : public std::regex {
public:
    regex() = default;
    regex(const char *p) : std::regex(refshim_translate(p)) {}
    regex(const std::string &p) : std::regex(refshim_translate(p)) {}
};
using std::smatch;
using std::sregex_token_iterator;
using std::regex_match;
using std::regex_search;
using std::regex_replace;
}  // namespace boost
