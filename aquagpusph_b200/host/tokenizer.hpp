// tokenizer.hpp -- scalar expression evaluator of the problem-definition files.
//
// Replaces aquagpusph/Tokenizer/* (muParser in the default build,
// CMakeLists.txt:126; ExprTk otherwise): expressions are evaluated in double
// precision over the registered scalar variables and then narrowed to the
// variable type by the caller (Tokenizer_muparser.hpp solve<T>, narrow_cast).
// Supported grammar (what the shipped presets and examples use):
//   numbers (1, 1.5, 1.e-8, 2.f), identifiers, function calls, ( ),
//   unary + - !, ^ (right assoc., binds tighter than unary minus), * / %,
//   + -, < > <= >= == !=, && ||, c ? a : b.
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace Aqua {

class Tokenizer {
  public:
    Tokenizer()
    {
        _vars["pi"] = M_PI;
        _vars["_pi"] = M_PI;
        _vars["e"] = M_E;
        _vars["_e"] = M_E;
        _vars["INFINITY"] = INFINITY;
    }

    /// @return true if the variable already existed
    bool registerVariable(const std::string& name, double value)
    {
        const bool had = _vars.count(name) > 0;
        _vars[name] = value;
        return had;
    }
    bool isVariable(const std::string& name) const { return _vars.count(name) > 0; }
    double variable(const std::string& name) const
    {
        auto it = _vars.find(name);
        return it == _vars.end() ? 0.0 : it->second;
    }

    /// Names of the variables an expression reads
    /// (all = true: every identifier that is not a function call or a constant,
    /// registered or not, so that the caller can reject unknown names at setup)
    std::vector<std::string> exprVariables(const std::string& eq, bool all = false) const
    {
        std::vector<std::string> out;
        std::set<std::string> seen;
        size_t i = 0;
        while (i < eq.size()) {
            if (isalpha((unsigned char)eq[i]) || eq[i] == '_') {
                size_t j = i;
                while (j < eq.size() && (isalnum((unsigned char)eq[j]) || eq[j] == '_'))
                    j++;
                const std::string id = eq.substr(i, j - i);
                size_t k = j;
                while (k < eq.size() && isspace((unsigned char)eq[k]))
                    k++;
                const bool is_call = k < eq.size() && eq[k] == '(';
                if (!is_call && (all || _vars.count(id)) && !seen.count(id) && id != "pi" && id != "e" &&
                    id != "_pi" && id != "_e" && id != "INFINITY") {
                    seen.insert(id);
                    out.push_back(id);
                }
                i = j;
            } else if (isdigit((unsigned char)eq[i]) || eq[i] == '.') {
                char* end;
                strtod(eq.c_str() + i, &end);
                size_t j = end - eq.c_str();
                if (j == i)
                    j = i + 1;
                if (j < eq.size() && (eq[j] == 'f' || eq[j] == 'F'))
                    j++;
                i = j;
            } else {
                i++;
            }
        }
        return out;
    }

    double solve(const std::string& eq) const
    {
        P p{ eq, 0, this };
        p.skip();
        if (p.i >= eq.size())
            throw std::runtime_error("Empty expression");
        const double v = p.ternary();
        p.skip();
        if (p.i != eq.size())
            throw std::runtime_error("Error evaluating \"" + eq + "\": unexpected token at " +
                                     std::to_string(p.i));
        return v;
    }

  private:
    std::map<std::string, double> _vars;

    struct P {
        const std::string& s;
        size_t i;
        const Tokenizer* t;

        void skip()
        {
            while (i < s.size() && isspace((unsigned char)s[i]))
                i++;
        }
        bool eat(const char* tok)
        {
            skip();
            const size_t n = strlen(tok);
            if (s.compare(i, n, tok))
                return false;
            i += n;
            return true;
        }
        [[noreturn]] void fail(const std::string& why)
        {
            throw std::runtime_error("Error evaluating \"" + s + "\": " + why + " at position " +
                                     std::to_string(i));
        }
        double ternary()
        {
            const double c = logic_or();
            skip();
            if (i < s.size() && s[i] == '?') {
                i++;
                const double a = ternary();
                if (!eat(":"))
                    fail("':' expected");
                const double b = ternary();
                return c != 0.0 ? a : b;
            }
            return c;
        }
        double logic_or()
        {
            double a = logic_and();
            while (eat("||")) {
                const double b = logic_and();
                a = (a != 0.0 || b != 0.0) ? 1.0 : 0.0;
            }
            return a;
        }
        double logic_and()
        {
            double a = compare();
            while (eat("&&")) {
                const double b = compare();
                a = (a != 0.0 && b != 0.0) ? 1.0 : 0.0;
            }
            return a;
        }
        double compare()
        {
            double a = additive();
            for (;;) {
                if (eat("<=")) a = a <= additive() ? 1.0 : 0.0;
                else if (eat(">=")) a = a >= additive() ? 1.0 : 0.0;
                else if (eat("==")) a = a == additive() ? 1.0 : 0.0;
                else if (eat("!=")) a = a != additive() ? 1.0 : 0.0;
                else if (eat("<")) a = a < additive() ? 1.0 : 0.0;
                else if (eat(">")) a = a > additive() ? 1.0 : 0.0;
                else return a;
            }
        }
        double additive()
        {
            double a = term();
            for (;;) {
                if (eat("+")) a += term();
                else if (eat("-")) a -= term();
                else return a;
            }
        }
        double term()
        {
            double a = unary();
            for (;;) {
                if (eat("*")) a *= unary();
                else if (eat("/")) a /= unary();
                else if (eat("%")) a = fmod(a, unary());
                else return a;
            }
        }
        double unary()
        {
            skip();
            if (i < s.size() && s[i] == '-') { i++; return -unary(); }
            if (i < s.size() && s[i] == '+') { i++; return unary(); }
            if (i < s.size() && s[i] == '!' && (i + 1 >= s.size() || s[i + 1] != '=')) {
                i++;
                return unary() == 0.0 ? 1.0 : 0.0;
            }
            return power();
        }
        double power()
        {
            const double b = primary();
            skip();
            if (i < s.size() && s[i] == '^') {
                i++;
                return pow(b, unary()); // right associative
            }
            return b;
        }
        double primary()
        {
            skip();
            if (i >= s.size())
                fail("unexpected end");
            if (s[i] == '(') {
                i++;
                const double v = ternary();
                if (!eat(")"))
                    fail("')' expected");
                return v;
            }
            if (isdigit((unsigned char)s[i]) || s[i] == '.') {
                char* end;
                const double v = strtod(s.c_str() + i, &end);
                const size_t j = end - s.c_str();
                if (j == i)
                    fail("bad number");
                i = j;
                if (i < s.size() && (s[i] == 'f' || s[i] == 'F'))
                    i++; // OpenCL style literal "2.f"
                return v;
            }
            if (isalpha((unsigned char)s[i]) || s[i] == '_') {
                size_t j = i;
                while (j < s.size() && (isalnum((unsigned char)s[j]) || s[j] == '_'))
                    j++;
                const std::string id = s.substr(i, j - i);
                i = j;
                skip();
                if (i < s.size() && s[i] == '(') {
                    i++;
                    std::vector<double> a;
                    skip();
                    if (i < s.size() && s[i] == ')') {
                        i++;
                    } else {
                        for (;;) {
                            a.push_back(ternary());
                            if (eat(","))
                                continue;
                            if (eat(")"))
                                break;
                            fail("',' or ')' expected");
                        }
                    }
                    return call(id, a);
                }
                auto it = t->_vars.find(id);
                if (it == t->_vars.end())
                    fail("unknown variable \"" + id + "\"");
                return it->second;
            }
            fail(std::string("unexpected character '") + s[i] + "'");
        }
        double call(const std::string& f, const std::vector<double>& a)
        {
            auto need = [&](size_t n) {
                if (a.size() != n)
                    fail("function " + f + " expects " + std::to_string(n) + " arguments");
            };
            if (f == "sqrt") { need(1); return sqrt(a[0]); }
            if (f == "abs") { need(1); return fabs(a[0]); }
            if (f == "sin") { need(1); return sin(a[0]); }
            if (f == "cos") { need(1); return cos(a[0]); }
            if (f == "tan") { need(1); return tan(a[0]); }
            if (f == "asin") { need(1); return asin(a[0]); }
            if (f == "acos") { need(1); return acos(a[0]); }
            if (f == "atan") { need(1); return atan(a[0]); }
            if (f == "atan2") { need(2); return atan2(a[0], a[1]); }
            if (f == "sinh") { need(1); return sinh(a[0]); }
            if (f == "cosh") { need(1); return cosh(a[0]); }
            if (f == "tanh") { need(1); return tanh(a[0]); }
            if (f == "exp") { need(1); return exp(a[0]); }
            if (f == "log" || f == "ln") { need(1); return std::log(a[0]); }
            if (f == "log2") { need(1); return log2(a[0]); }
            if (f == "log10") { need(1); return log10(a[0]); }
            if (f == "floor") { need(1); return floor(a[0]); }
            if (f == "ceil") { need(1); return ceil(a[0]); }
            if (f == "round" || f == "rint") { need(1); return rint(a[0]); }
            if (f == "sign") { need(1); return (a[0] > 0) - (a[0] < 0); }
            if (f == "pow") { need(2); return pow(a[0], a[1]); }
            if (f == "min" || f == "max" || f == "sum" || f == "avg") {
                if (a.empty())
                    fail("function " + f + " needs arguments");
                double r = a[0];
                for (size_t k = 1; k < a.size(); k++)
                    r = f == "min" ? fmin(r, a[k]) : (f == "max" ? fmax(r, a[k]) : r + a[k]);
                return f == "avg" ? r / a.size() : r;
            }
            fail("unknown function \"" + f + "\"");
        }
    };
};

} // namespace Aqua
