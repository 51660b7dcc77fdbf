// xml.hpp -- a small, lenient XML reader for the problem-definition files.
//
// The reference parses with Xerces-C without installing an error handler
// (InputOutput/State.cpp:296-303), so files that are not well formed still yield
// the partial DOM built so far; five shipped presets depend on that
// (resources/Presets/src/basic/kernels/*.xml lack </Variables>).  This reader
// reproduces the leniency: a closing tag that does not match the open element
// closes every element up to the matching ancestor, and end-of-file closes
// whatever is still open.
#pragma once
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace Aqua {
namespace Xml {

struct Node {
    std::string tag;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::string text; // concatenated character data (getTextContent)
    std::vector<std::unique_ptr<Node>> children;
    Node* parent = nullptr;

    bool has(const std::string& a) const
    {
        for (auto& kv : attrs)
            if (kv.first == a)
                return true;
        return false;
    }
    std::string attr(const std::string& a) const
    {
        for (auto& kv : attrs)
            if (kv.first == a)
                return kv.second;
        return "";
    }
    // getElementsByTagName: every descendant with that tag, in document order
    void descendants(const std::string& t, std::vector<const Node*>& out) const
    {
        for (auto& c : children) {
            if (c->tag == t)
                out.push_back(c.get());
            c->descendants(t, out);
        }
    }
    std::vector<const Node*> descendants(const std::string& t) const
    {
        std::vector<const Node*> out;
        descendants(t, out);
        return out;
    }
};

inline std::string decodeEntities(const std::string& s)
{
    std::string o;
    o.reserve(s.size());
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] != '&') {
            o.push_back(s[i]);
            continue;
        }
        const size_t e = s.find(';', i);
        if (e == std::string::npos) {
            o.push_back(s[i]);
            continue;
        }
        const std::string ent = s.substr(i + 1, e - i - 1);
        if (ent == "lt") o.push_back('<');
        else if (ent == "gt") o.push_back('>');
        else if (ent == "amp") o.push_back('&');
        else if (ent == "quot") o.push_back('"');
        else if (ent == "apos") o.push_back('\'');
        else if (!ent.empty() && ent[0] == '#') {
            const long c = (ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X'))
                               ? strtol(ent.c_str() + 2, nullptr, 16)
                               : strtol(ent.c_str() + 1, nullptr, 10);
            o.push_back((char)c);
        } else {
            o += s.substr(i, e - i + 1);
        }
        i = e;
    }
    return o;
}

inline std::string encodeEntities(const std::string& s)
{
    std::string o;
    for (char c : s) {
        switch (c) {
            case '<': o += "&lt;"; break;
            case '>': o += "&gt;"; break;
            case '&': o += "&amp;"; break;
            case '"': o += "&quot;"; break;
            default: o.push_back(c);
        }
    }
    return o;
}

inline std::unique_ptr<Node> parseString(const std::string& s)
{
    auto doc = std::make_unique<Node>();
    doc->tag = "#document";
    Node* cur = doc.get();
    size_t i = 0;
    const size_t n = s.size();
    while (i < n) {
        if (s[i] != '<') {
            const size_t e = s.find('<', i);
            const size_t stop = (e == std::string::npos) ? n : e;
            cur->text += decodeEntities(s.substr(i, stop - i));
            i = stop;
            continue;
        }
        if (!s.compare(i, 4, "<!--")) {
            const size_t e = s.find("-->", i + 4);
            i = (e == std::string::npos) ? n : e + 3;
            continue;
        }
        if (!s.compare(i, 2, "<?")) {
            const size_t e = s.find("?>", i + 2);
            i = (e == std::string::npos) ? n : e + 2;
            continue;
        }
        if (!s.compare(i, 9, "<![CDATA[")) {
            const size_t e = s.find("]]>", i + 9);
            const size_t stop = (e == std::string::npos) ? n : e;
            cur->text += s.substr(i + 9, stop - i - 9);
            i = (e == std::string::npos) ? n : e + 3;
            continue;
        }
        if (!s.compare(i, 2, "<!")) { // DOCTYPE etc.
            const size_t e = s.find('>', i);
            i = (e == std::string::npos) ? n : e + 1;
            continue;
        }
        if (!s.compare(i, 2, "</")) {
            const size_t e = s.find('>', i);
            const std::string tag = s.substr(i + 2, (e == std::string::npos ? n : e) - i - 2);
            std::string t;
            for (char c : tag)
                if (!isspace((unsigned char)c))
                    t.push_back(c);
            // close up to the matching ancestor (lenient); ignore if none
            Node* a = cur;
            while (a && a->tag != t)
                a = a->parent;
            if (a && a->parent)
                cur = a->parent;
            i = (e == std::string::npos) ? n : e + 1;
            continue;
        }
        // opening tag
        size_t j = i + 1;
        while (j < n && !isspace((unsigned char)s[j]) && s[j] != '>' && s[j] != '/')
            j++;
        auto node = std::make_unique<Node>();
        node->tag = s.substr(i + 1, j - i - 1);
        node->parent = cur;
        bool selfclose = false;
        while (j < n) {
            while (j < n && isspace((unsigned char)s[j]))
                j++;
            if (j >= n)
                break;
            if (s[j] == '>') {
                j++;
                break;
            }
            if (s[j] == '/') {
                selfclose = true;
                j++;
                continue;
            }
            size_t k = j;
            while (k < n && s[k] != '=' && !isspace((unsigned char)s[k]) && s[k] != '>' && s[k] != '/')
                k++;
            const std::string name = s.substr(j, k - j);
            while (k < n && isspace((unsigned char)s[k]))
                k++;
            std::string value;
            if (k < n && s[k] == '=') {
                k++;
                while (k < n && isspace((unsigned char)s[k]))
                    k++;
                if (k < n && (s[k] == '"' || s[k] == '\'')) {
                    const char q = s[k];
                    const size_t e = s.find(q, k + 1);
                    const size_t stop = (e == std::string::npos) ? n : e;
                    value = decodeEntities(s.substr(k + 1, stop - k - 1));
                    k = (e == std::string::npos) ? n : e + 1;
                }
            }
            if (!name.empty())
                node->attrs.emplace_back(name, value);
            j = k;
        }
        Node* raw = node.get();
        cur->children.push_back(std::move(node));
        if (!selfclose)
            cur = raw;
        i = j;
    }
    return doc;
}

inline std::unique_ptr<Node> parseFile(const std::string& path)
{
    std::ifstream f(path);
    if (!f)
        throw std::runtime_error("File inaccessible: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parseString(ss.str());
}

// The document element (first element child), nullptr when the file is empty
inline const Node* root(const Node* doc)
{
    for (auto& c : doc->children)
        return c.get();
    return nullptr;
}

} // namespace Xml
} // namespace Aqua
