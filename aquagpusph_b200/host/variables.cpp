// variables.cpp -- see variables.hpp
#include "variables.hpp"

#include <cmath>
#include <limits>

namespace Aqua {

int& logLevel()
{
    static int lvl = L_WARNING;
    return lvl;
}

void log(LogLevel l, const std::string& msg)
{
    if ((int)l < logLevel())
        return;
    static const char* tag[] = { "DEBUG", "INFO", "WARNING", "ERROR" };
    fprintf(stderr, "[%s] %s%s", tag[l], msg.c_str(),
            (!msg.empty() && msg.back() == '\n') ? "" : "\n");
}

namespace InputOutput {

static const char* const EXT[16] = { "_x",  "_y",  "_z",  "_w",  "_yx", "_yy", "_yz", "_yw",
                                     "_zx", "_zy", "_zz", "_zw", "_wx", "_wy", "_wz", "_ww" };

std::string Variable::asString() const
{
    std::ostringstream s;
    if (isArray()) {
        s << _dptr;
        return s.str();
    }
    if (_ncomp > 1)
        s << "(";
    for (unsigned c = 0; c < _ncomp; c++) {
        const char* p = _value.data() + c * compsize();
        if (c)
            s << ",";
        char b[64];
        switch (_kind) {
            case 'i': s << *(const int32_t*)p; break;
            case 'u': s << *(const uint32_t*)p; break;
            case 'l': s << *(const int64_t*)p; break;
            case 'L': s << *(const uint64_t*)p; break;
            case 'd': snprintf(b, sizeof(b), "%.16g", *(const double*)p); s << b; break;
            default: snprintf(b, sizeof(b), "%.9g", (double)*(const float*)p); s << b;
        }
    }
    if (_ncomp > 1)
        s << ")";
    return s.str();
}

Variables::~Variables()
{
    for (auto& v : _vars)
        if (v->isArray() && v->_dptr)
            aqc_free(_ctx, v->_dptr);
}

std::string Variables::typeAlias(const std::string& t) const
{
    const bool d3 = _dims == 3;
    if (t == "int32") return "int";
    if (t == "int64") return "long";
    if (t == "uint32" || t == "uint") return "unsigned int";
    if (t == "uint64") return "unsigned long";
    if (t == "size_t" || t == "usize") return "unsigned int"; // 32-bit device addressing (State.cpp:499-502)
    if (t == "ssize_t") return "int";
    if (startswith(t, "svec")) return typeAlias(replaceAllCopy(t, "svec", "uivec"));
    if (startswith(t, "ssvec")) return typeAlias(replaceAllCopy(t, "ssvec", "ivec"));
    if (startswith(t, "int") && t != "int") return replaceAllCopy(t, "int", "ivec");
    if (startswith(t, "long") && t != "long") return replaceAllCopy(t, "long", "lvec");
    if (startswith(t, "uint") && t != "uint") return replaceAllCopy(t, "uint", "uivec");
    if (startswith(t, "ulong") && t != "ulong") return replaceAllCopy(t, "ulong", "ulvec");
    if (startswith(t, "float") && t != "float") return replaceAllCopy(t, "float", "vec");
    if (startswith(t, "fvec") && t != "fvec") return replaceAllCopy(t, "fvec", "vec");
    if (startswith(t, "double") && t != "double") return replaceAllCopy(t, "double", "dvec");
    if (t == "ivec") return d3 ? "ivec4" : "ivec2";
    if (t == "lvec") return d3 ? "lvec4" : "lvec2";
    if (t == "uivec") return d3 ? "uivec4" : "uivec2";
    if (t == "ulvec") return d3 ? "ulvec4" : "ulvec2";
    if (t == "vec" || t == "fvec") return d3 ? "vec4" : "vec2";
    if (t == "dvec") return d3 ? "dvec4" : "dvec2";
    return t;
}

unsigned Variables::typeToN(const std::string& type) const
{
    if (type.find("vec2") != std::string::npos) return 2;
    if (type.find("vec3") != std::string::npos) return 3;
    if (type.find("vec4") != std::string::npos) return 4;
    if (type.find("vec8") != std::string::npos) return 8;
    if (type.find("vec") != std::string::npos) return _dims == 3 ? 4 : 2;
    if (type.find("matrix") != std::string::npos) return _dims == 3 ? 16 : 4;
    return 1;
}

void Variables::describe(const std::string& type_in, size_t& typesize, unsigned& ncomp,
                         char& kind) const
{
    std::string t = trimCopy(type_in);
    if (!t.empty() && t.back() == '*')
        t.pop_back();
    t = typeAlias(trimCopy(t));
    ncomp = typeToN(t);
    if (ncomp == 3)
        ncomp = 4; // 3-component OpenCL vectors are stored as 4 (Variable.cpp:1274)
    size_t cs = 0;
    if (t.find("unsigned int") != std::string::npos || t.find("uivec") != std::string::npos) {
        cs = 4; kind = 'u';
    } else if (t.find("unsigned long") != std::string::npos || t.find("ulvec") != std::string::npos) {
        cs = 8; kind = 'L';
    } else if (t.find("int") != std::string::npos || t.find("ivec") != std::string::npos) {
        cs = 4; kind = 'i';
    } else if (t.find("long") != std::string::npos || t.find("lvec") != std::string::npos) {
        cs = 8; kind = 'l';
    } else if (t.find("double") != std::string::npos || t.find("dvec") != std::string::npos) {
        cs = 8; kind = 'd';
    } else if (t.find("float") != std::string::npos || t.find("vec") != std::string::npos ||
               t.find("matrix") != std::string::npos) {
        cs = 4; kind = 'f';
    } else {
        cs = 0; kind = '?';
    }
    typesize = cs * ncomp;
}

size_t Variables::typeToBytes(const std::string& type) const
{
    size_t ts;
    unsigned n;
    char k;
    describe(type, ts, n, k);
    return ts;
}

bool Variables::isSameType(const std::string& a, const std::string& b, bool ignore_asterisk) const
{
    if (!ignore_asterisk) {
        const bool pa = a.find('*') != std::string::npos, pb = b.find('*') != std::string::npos;
        if (pa != pb)
            return false;
    }
    size_t sa, sb;
    unsigned na, nb;
    char ka, kb;
    describe(a, sa, na, ka);
    describe(b, sb, nb, kb);
    return sa == sb && na == nb && ka == kb;
}

Variable* Variables::get(const std::string& name) const
{
    for (auto& v : _vars)
        if (v->name() == name)
            return v.get();
    return nullptr;
}

void Variables::registerVariable(const std::string& name, const std::string& type,
                                 const std::string& length, const std::string& value)
{
    // an already existing variable with the same name is replaced (Variable.cpp:1058-1064)
    for (size_t i = 0; i < _vars.size(); i++)
        if (_vars[i]->name() == name) {
            if (_vars[i]->isArray() && _vars[i]->_dptr)
                aqc_free(_ctx, _vars[i]->_dptr);
            _vars.erase(_vars.begin() + i);
            break;
        }
    auto v = std::make_unique<Variable>(name, trimCopy(type));
    describe(type, v->_typesize, v->_ncomp, v->_kind);
    if (!v->_typesize)
        throw std::runtime_error("Invalid type \"" + type + "\" for variable \"" + name + "\"");
    if (v->isArray()) {
        uint64_t n = 0;
        if (!trimCopy(length).empty()) {
            try {
                solve("unsigned long", length, &n);
            } catch (std::exception& e) {
                throw std::runtime_error("Invalid array length \"" + length + "\" for variable \"" +
                                         name + "\": " + e.what());
            }
        }
        v->_length = (size_t)n;
        if (n) {
            void* p = nullptr;
            if (aqc_alloc(_ctx, n * v->_typesize, &p))
                throw std::runtime_error(std::string("Failure allocating \"") + name +
                                         "\": " + aqc_last_error(_ctx));
            v->_dptr = p;
        }
        _vars.push_back(std::move(v));
        return;
    }
    v->_value.assign(v->_typesize, 0);
    Variable* raw = v.get();
    _vars.push_back(std::move(v));
    if (!trimCopy(value).empty())
        solve(raw->type(), value, raw->get(), name);
    else
        populate(raw);
}

static void store(char kind, void* dst, double v)
{
    // narrow_cast<T> (boost::numeric_cast): truncation with a range check
    auto range = [&](double lo, double hi) {
        if (std::isnan(v) || v < lo || v > hi)
            throw std::out_of_range("value " + std::to_string(v) + " overflows the variable type");
    };
    switch (kind) {
        case 'i': range(-2147483648.0, 2147483647.0); *(int32_t*)dst = (int32_t)v; break;
        case 'u': range(0.0, 4294967295.0); *(uint32_t*)dst = (uint32_t)v; break;
        case 'l': *(int64_t*)dst = (int64_t)v; break;
        case 'L': range(0.0, 1.8446744073709552e19); *(uint64_t*)dst = (uint64_t)v; break;
        case 'd': *(double*)dst = v; break;
        default: *(float*)dst = (float)v;
    }
}

static double load(char kind, const void* src)
{
    switch (kind) {
        case 'i': return *(const int32_t*)src;
        case 'u': return *(const uint32_t*)src;
        case 'l': return (double)*(const int64_t*)src;
        case 'L': return (double)*(const uint64_t*)src;
        case 'd': return *(const double*)src;
        default: return *(const float*)src;
    }
}

void Variables::solve(const std::string& type, const std::string& expr, void* data,
                      const std::string& name)
{
    size_t ts;
    unsigned n;
    char kind;
    describe(type, ts, n, kind);
    if (!ts)
        throw std::runtime_error("0 bytes size for type \"" + type + "\"");
    if (trimCopy(expr).empty())
        throw std::runtime_error("Empty expression");
    const std::vector<std::string> parts = split_formulae(expr);
    if (parts.size() < n)
        throw std::runtime_error("Invalid number of fields in \"" + expr + "\" (" +
                                 std::to_string(n) + " expected)");
    const size_t cs = ts / n;
    for (unsigned c = 0; c < n; c++) {
        const double v = tok.solve(parts[c]);
        store(kind, (char*)data + c * cs, v);
        if (!name.empty()) {
            // the tokenizer sees the narrowed value (Variable.cpp:1253-1292)
            const double back = load(kind, (char*)data + c * cs);
            tok.registerVariable(n == 1 ? name : name + EXT[c], back);
        }
    }
}

void Variables::populate(Variable* var)
{
    if (!var || var->isArray())
        return;
    const unsigned n = var->ncomp();
    for (unsigned c = 0; c < n && c < 16; c++) {
        const double v = load(var->kind(), (const char*)var->get() + c * var->compsize());
        tok.registerVariable(n == 1 ? var->name() : var->name() + EXT[c], v);
    }
}

void Variables::populate(const std::string& name) { populate(get(name)); }

std::vector<Variable*> Variables::exprVariables(const std::string& expr) const
{
    std::vector<Variable*> out;
    for (auto part : split_formulae(expr))
        for (auto id : tok.exprVariables(part, true)) {
            std::string vn = id;
            for (auto suf : { "_x", "_y", "_z", "_w" })
                if (endswith(vn, suf) && !get(vn)) {
                    vn.erase(vn.size() - 2);
                    break;
                }
            Variable* v = get(vn);
            // Variable.cpp:1230-1237: an expression naming an unknown variable is an error
            if (!v && !tok.isVariable(id))
                throw std::runtime_error("Variable \"" + id + "\", referenced on the expression " +
                                         expr + ", cannot be found");
            if (v && std::find(out.begin(), out.end(), v) == out.end())
                out.push_back(v);
        }
    return out;
}

size_t Variables::allocatedMemory() const
{
    size_t m = 0;
    for (auto& v : _vars)
        if (v->isArray())
            m += v->size();
    return m;
}

} // namespace InputOutput
} // namespace Aqua
