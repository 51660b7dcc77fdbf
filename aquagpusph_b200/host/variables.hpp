// variables.hpp -- typed scalar and device-array variables of a simulation and
// their registry.  Counterpart of aquagpusph/Variable.{hpp,cpp}: same type names
// and aliases (Variable.cpp:1547-1624), same scalar/array split, same tokenizer
// population (vector components are published as name_x .. name_w,
// Variable.cpp:1208-1292).  Arrays live in device memory owned through the C-ABI
// (aqc_alloc), scalars on the host.
#pragma once
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "aquacuda.h"
#include "aux.hpp"
#include "tokenizer.hpp"

namespace Aqua {
namespace InputOutput {

class Variables;

class Variable {
  public:
    Variable(const std::string& name, const std::string& type) : _name(name), _type(type) {}
    const std::string& name() const { return _name; }
    const std::string& type() const { return _type; }
    bool isArray() const { return _type.find('*') != std::string::npos; }
    bool isScalar() const { return !isArray(); }

    // --- scalars: raw bytes in the variable's type
    void* get() { return _value.data(); }
    const void* get() const { return _value.data(); }
    size_t typesize() const { return _typesize; }
    void set(const void* data) { memcpy(_value.data(), data, _typesize); }

    // --- arrays
    void* dptr() const { return _dptr; }
    size_t length() const { return _length; }   // elements
    size_t size() const { return isArray() ? _length * _typesize : _typesize; } // bytes
    bool reallocatable() const { return _realloc; }
    void reallocatable(bool v) { _realloc = v; }
    /// Replace the device buffer (LinkList growing ihoc, LinkList.cpp:249-270)
    void reset(void* dptr, size_t length) { _dptr = dptr; _length = length; }

    char kind() const { return _kind; }     // 'i' int32, 'u' uint32, 'l' int64, 'L' uint64, 'f' float, 'd' double
    unsigned ncomp() const { return _ncomp; }
    size_t compsize() const { return _typesize / _ncomp; }
    std::string asString() const;

  private:
    friend class Variables;
    std::string _name, _type;
    std::vector<char> _value;
    size_t _typesize = 0;
    unsigned _ncomp = 1;
    char _kind = 'f';
    void* _dptr = nullptr;
    size_t _length = 0;
    bool _realloc = false;
};

class Variables {
  public:
    Variables(int dims, aqc_ctx* ctx) : _dims(dims), _ctx(ctx) {}
    ~Variables();

    void registerVariable(const std::string& name, const std::string& type,
                          const std::string& length, const std::string& value);
    Variable* get(const std::string& name) const;
    const std::vector<std::unique_ptr<Variable>>& all() const { return _vars; }

    /// bytes of one element of `type` (0 when unknown); '*' is ignored
    size_t typeToBytes(const std::string& type) const;
    unsigned typeToN(const std::string& type) const;
    std::string typeAlias(const std::string& t) const;
    bool isSameType(const std::string& a, const std::string& b, bool ignore_asterisk = true) const;

    /// Evaluate `expr` as a value of `type` into `data`; when `name` is given the
    /// result is also published in the tokenizer (Variable.cpp:1321-1435).
    void solve(const std::string& type, const std::string& expr, void* data,
               const std::string& name = "");
    /// Publish the current value of a scalar variable in the tokenizer
    void populate(Variable* var);
    void populate(const std::string& name);
    /// Variables an expression depends on (Variable.cpp:1221-1238)
    std::vector<Variable*> exprVariables(const std::string& expr) const;

    Tokenizer& tokenizer() { return tok; }
    size_t allocatedMemory() const;
    int dims() const { return _dims; }
    aqc_ctx* ctx() const { return _ctx; }

  private:
    void describe(const std::string& type, size_t& typesize, unsigned& ncomp, char& kind) const;
    int _dims;
    aqc_ctx* _ctx;
    Tokenizer tok;
    std::vector<std::unique_ptr<Variable>> _vars;
};

} // namespace InputOutput
} // namespace Aqua
