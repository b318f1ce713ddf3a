// seed_text.h — the SEED function-string helpers of the reference
// (src/seed_utils.h) without Boost.Regex: each regex is matched by hand with
// the same leftmost / lazy / greedy outcome.
#pragma once

#include <string>
#include <vector>

namespace sigk_host {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

// split_func_comment: regex_match of "(.*?)(?:\s+(\#+)\s+(.*))?"  (src/seed_utils.h:13,30-44).
// The lazy first group stops at the first position where  \s+ #+ \s+  matches.
inline void split_func_comment(const std::string &str, std::string &func, std::string &sep, std::string &comment) {
    const size_t n = str.size();
    for (size_t i = 0; i < n; ++i) {
        if (!is_space(str[i])) continue;
        size_t j = i;
        while (j < n && is_space(str[j])) ++j;          // \s+ (greedy; the next atom is not a space, so no backtracking matters)
        size_t k = j;
        while (k < n && str[k] == '#') ++k;              // #+
        if (k == j) continue;
        size_t m = k;
        while (m < n && is_space(str[m])) ++m;          // \s+
        if (m == k) {
            // "#+" greedy took every '#'; with fewer it still needs a space after it: no match from i
            continue;
        }
        func = str.substr(0, i);
        sep = str.substr(j, k - j);
        comment = str.substr(m);
        return;
    }
    func = str; sep.clear(); comment.clear();
}

// is_truncated_comment: regex_search "^(?:frag|missing|trunc)"  (src/seed_utils.h:17,45-48)
inline bool is_truncated_comment(const std::string &s) {
    return s.compare(0, 4, "frag") == 0 || s.compare(0, 7, "missing") == 0 || s.compare(0, 5, "trunc") == 0;
}

// strip_func_comment: regex_replace "(\s*\#.*$)" -> ""  (src/seed_utils.h:12,25-29)
inline std::string strip_func_comment(const std::string &s) {
    const size_t h = s.find('#');
    if (h == std::string::npos) return s;
    size_t b = h;
    while (b > 0 && is_space(s[b - 1])) --b;
    return s.substr(0, b);
}

// roles_of_function: split on "\s+[/@]\s+|\s*;\s+"  (src/seed_utils.h:15,50-62).
// sregex_token_iterator(-1): pieces between matches; a leading empty piece is kept, a trailing one is not.
inline std::vector<std::string> roles_of_function(const std::string &function) {
    const std::string s = strip_func_comment(function);
    const size_t n = s.size();
    std::vector<std::string> out;
    size_t piece = 0, i = 0;
    auto match_at = [&](size_t p) -> size_t {        // end of a delimiter starting at p, or 0
        size_t j = p;
        while (j < n && is_space(s[j])) ++j;
        if (j > p && j < n && (s[j] == '/' || s[j] == '@')) {       // \s+[/@]\s+
            size_t k = j + 1;
            while (k < n && is_space(s[k])) ++k;
            if (k > j + 1) return k;
        }
        if (j < n && s[j] == ';') {                                 // \s*;\s+
            size_t k = j + 1;
            while (k < n && is_space(s[k])) ++k;
            if (k > j + 1) return k;
        }
        return 0;
    };
    while (i < n) {
        const size_t e = match_at(i);
        if (e) { out.emplace_back(s.substr(piece, i - piece)); piece = i = e; }
        else ++i;
    }
    if (piece < n || out.empty()) out.emplace_back(s.substr(piece));
    return out;
}

}  // namespace sigk_host
