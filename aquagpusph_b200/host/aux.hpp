// aux.hpp -- small helpers of the C++ host (string utilities, rounding, wildcard
// matching).  Behavioural counterparts: aquagpusph/AuxiliarMethods.hpp:258-305
// (nextPowerOf2, roundUp) and InputOutput/State.cpp:58-89 (wildcard match).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace Aqua {

inline std::string trimCopy(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a]))
        a++;
    while (b > a && isspace((unsigned char)s[b - 1]))
        b--;
    return s.substr(a, b - a);
}

inline std::string toLowerCopy(std::string s)
{
    for (auto& c : s)
        c = (char)tolower((unsigned char)c);
    return s;
}

inline bool startswith(const std::string& s, const std::string& p)
{
    return s.size() >= p.size() && !s.compare(0, p.size(), p);
}

inline bool endswith(const std::string& s, const std::string& p)
{
    return s.size() >= p.size() && !s.compare(s.size() - p.size(), p.size(), p);
}

inline std::string replaceAllCopy(std::string s, const std::string& from, const std::string& to)
{
    if (from.empty())
        return s;
    size_t pos = 0;
    while ((pos = s.find(from, pos)) != std::string::npos) {
        s.replace(pos, from.size(), to);
        pos += to.size();
    }
    return s;
}

inline std::vector<std::string> split(const std::string& s, char sep = ',')
{
    std::vector<std::string> out;
    std::string item;
    std::istringstream f(s);
    while (std::getline(f, item, sep))
        out.push_back(trimCopy(item));
    return out;
}

// Split "a, f(b, c), d" at top-level commas (vector-valued expressions).
inline std::vector<std::string> split_formulae(const std::string& s)
{
    std::vector<std::string> out;
    int depth = 0;
    std::string cur;
    for (char c : s) {
        if (c == '(')
            depth++;
        else if (c == ')')
            depth--;
        if ((c == ',' || c == ';') && depth == 0) {
            out.push_back(trimCopy(cur));
            cur.clear();
        } else {
            cur.push_back(c);
        }
    }
    if (!trimCopy(cur).empty())
        out.push_back(trimCopy(cur));
    return out;
}

template <typename T>
inline bool isPowerOf2(T x)
{
    return !(x & (x - 1));
}

template <typename T>
inline T nextPowerOf2(T n)
{
    if (n && isPowerOf2(n))
        return n;
    T p = 1;
    while (p < n)
        p <<= 1;
    return p;
}

template <typename T>
inline T roundUp(T x, T divisor)
{
    T rest = x % divisor;
    if (rest) {
        x -= rest;
        x += divisor;
    }
    return x;
}

// '*' matches any run of characters (including none)
inline bool wildcardMatch(const char* p, const char* s)
{
    if (!*p && !*s)
        return true;
    if (*p == '*' && *(p + 1) && !*s)
        return false;
    if (*p == *s && *p)
        return wildcardMatch(p + 1, s + 1);
    if (*p == '*')
        return wildcardMatch(p + 1, s) || (*s && wildcardMatch(p, s + 1));
    return false;
}

inline bool match(const std::string& pattern, const std::string& name)
{
    if (pattern == name)
        return true;
    return wildcardMatch(pattern.c_str(), name.c_str());
}

// Replace "{name}" placeholders in file names (AuxiliarMethods.hpp:172-255)
inline std::string formatPath(std::string s, int mpi_rank, int index = -1)
{
    s = replaceAllCopy(s, "{mpi_rank}", std::to_string(mpi_rank));
    if (index >= 0) {
        char b[32];
        snprintf(b, sizeof(b), "%05d", index);
        s = replaceAllCopy(s, "{index}", b);
    }
    return s;
}

enum LogLevel { L_DEBUG = 0, L_INFO = 1, L_WARNING = 2, L_ERROR = 3 };
int& logLevel();
void log(LogLevel l, const std::string& msg);

} // namespace Aqua
