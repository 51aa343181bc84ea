/*
 * oracle/ref_shim/exprtk.hpp -- TEST INFRASTRUCTURE (see oracle/tm_oracle.h).
 *
 * The reference includes ArashPartow/exprtk (<exprtk.hpp>, rietkerk_bjork_basic.h:28) but does not
 * vendor it and pins no version; there is no network here.  This is NOT exprtk: it is a
 * from-scratch stand-in exposing just the calls the reference makes (symbol_table::add_variable /
 * remove_variable / get_variable_list / symbol_exists / get_variable, expression::
 * register_symbol_table / value, parser::compile) and evaluating the only grammar its drivers
 * generate (order_1stoc_test_rietkerk.cpp:151-158): sums and products of numbers and variables,
 * e.g. "40000.000000 * a + -8000.000000".
 */
#ifndef TM_EXPRTK_SHIM_HPP
#define TM_EXPRTK_SHIM_HPP
#include <cstdlib>
#include <cctype>
#include <map>
#include <string>
#include <vector>

namespace exprtk {

template <typename T> class symbol_table {
public:
    bool add_variable(const std::string &name, T &ref) { vars_[name] = &ref; return true; }
    bool remove_variable(const std::string &name) { return vars_.erase(name) > 0; }
    bool add_constants() { return true; }
    bool symbol_exists(const std::string &name) const { return vars_.count(name) > 0; }
    std::size_t get_variable_list(std::vector<std::string> &out) const
    {
        for (auto &kv : vars_) out.push_back(kv.first);
        return vars_.size();
    }
    const T *get_variable(const std::string &name) const
    {
        auto it = vars_.find(name);
        return it == vars_.end() ? nullptr : it->second;
    }
private:
    std::map<std::string, T *> vars_;
};

template <typename T> class expression {
public:
    void register_symbol_table(symbol_table<T> &st) { st_ = &st; }
    T value() const { std::size_t p = 0; return sum(p); }
    void set_source(const std::string &s) { src_ = s; }
private:
    void ws(std::size_t &p) const { while (p < src_.size() && std::isspace((unsigned char)src_[p])) p++; }
    T atom(std::size_t &p) const
    {
        ws(p);
        if (p < src_.size() && src_[p] == '(') { p++; T v = sum(p); ws(p); if (p < src_.size()) p++; return v; }
        if (p < src_.size() && (src_[p] == '-' || src_[p] == '+')) { bool neg = src_[p] == '-'; p++; T v = atom(p); return neg ? -v : v; }
        if (p < src_.size() && (std::isalpha((unsigned char)src_[p]) || src_[p] == '_')) {
            std::size_t q = p;
            while (q < src_.size() && (std::isalnum((unsigned char)src_[q]) || src_[q] == '_')) q++;
            std::string name = src_.substr(p, q - p); p = q;
            const T *v = st_ ? st_->get_variable(name) : nullptr;
            return v ? *v : T(0);
        }
        const char *b = src_.c_str() + p; char *e = nullptr;
        T v = (T)std::strtod(b, &e); p += (std::size_t)(e - b);
        return v;
    }
    T prod(std::size_t &p) const
    {
        T v = atom(p);
        for (;;) { ws(p);
            if (p < src_.size() && src_[p] == '*') { p++; v = v * atom(p); }
            else if (p < src_.size() && src_[p] == '/') { p++; v = v / atom(p); }
            else return v; }
    }
    T sum(std::size_t &p) const
    {
        T v = prod(p);
        for (;;) { ws(p);
            if (p < src_.size() && src_[p] == '+') { p++; v = v + prod(p); }
            else if (p < src_.size() && src_[p] == '-') { p++; v = v - prod(p); }
            else return v; }
    }
    symbol_table<T> *st_ = nullptr;
    std::string src_;
};

template <typename T> class parser {
public:
    bool compile(const std::string &s, expression<T> &e) { e.set_source(s); return true; }
    std::string error() const { return ""; }
};

}
#endif
