// TEST INFRASTRUCTURE ONLY — stand-in for FSL's utils/options.h (Utilities::OptionParser), which is
// not installed and not part of /root/reference. It lets the UNMODIFIED reference sources
// (mesh_registration.cpp:459-760, src/msmOptions.h, src/newmsm.cpp) compile and parse the reference's
// own command lines and config files ("--key=v1,v2,..." one per line, '#' comments, bare "--flag").
// Own implementation of the public surface the reference names; no FSL code.
#ifndef ORACLE_SHIM_UTILS_OPTIONS_H
#define ORACLE_SHIM_UTILS_OPTIONS_H
#include <cstdlib>
#include <exception>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace Utilities {

typedef enum { no_argument = 0, requires_argument, optional_argument, requires_2_arguments } ArgFlag;
enum OverwriteMode { Allow = 0, ThrowException, Ignore };

class X_OptionError : public std::exception {
    std::string msg;
public:
    X_OptionError(const std::string& o, const std::string& e) : msg(o + ": " + e + "!") {}
    explicit X_OptionError(const std::string& e) : msg(e) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

class BaseOption {
protected:
    std::string key_, help_;
    ArgFlag flag_;
    bool compulsory_, set_ = false;
public:
    BaseOption(const std::string& k, const std::string& h, bool c, ArgFlag f) : key_(k), help_(h), flag_(f), compulsory_(c) {}
    virtual ~BaseOption() = default;
    bool set() const { return set_; }
    bool unset() const { return !set_; }
    bool compulsory() const { return compulsory_; }
    bool has_arg() const { return flag_ != no_argument; }
    const std::string& key() const { return key_; }
    const std::string& help_text() const { return help_; }
    // "-i,--inmesh" matches "-i" and "--inmesh"
    bool matches(const std::string& arg) const {
        std::stringstream ss(key_);
        std::string k;
        while (std::getline(ss, k, ',')) if (k == arg) return true;
        return false;
    }
    virtual bool set_value(const std::string& v) = 0;
    void mark_set() { set_ = true; }
};

namespace detail {
inline bool conv(const std::string& s, std::string& out) { out = s; return true; }
inline bool conv(const std::string& s, int& out) { char* e; out = (int)std::strtol(s.c_str(), &e, 10); return e != s.c_str(); }
inline bool conv(const std::string& s, float& out) { char* e; out = std::strtof(s.c_str(), &e); return e != s.c_str(); }
inline bool conv(const std::string& s, double& out) { char* e; out = std::strtod(s.c_str(), &e); return e != s.c_str(); }
inline bool conv(const std::string& s, bool& out) {
    if (s == "true" || s == "1" || s.empty()) { out = true; return true; }
    if (s == "false" || s == "0") { out = false; return true; }
    return false;
}
template <class T> inline bool conv(const std::string& s, std::vector<T>& out) {
    out.clear();
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) { T v; if (!conv(tok, v)) return false; out.push_back(v); }
    return true;
}
}  // namespace detail

template <class T> class Option : public BaseOption {
    T value_, default_;
public:
    Option(const std::string& k, const T& def, const std::string& h, bool c, ArgFlag f = no_argument, bool /*visible*/ = true)
        : BaseOption(k, h, c, f), value_(def), default_(def) {}
    const T& value() const { return value_; }
    const T& default_value() const { return default_; }
    bool set_value(const std::string& v) override {
        if (!has_arg()) { T t = value_; detail::conv(std::string("true"), t); value_ = t; set_ = true; return true; }
        T t;
        if (!detail::conv(v, t)) return false;
        value_ = t; set_ = true; return true;
    }
    bool set_T(const T& v) { value_ = v; set_ = true; return true; }
};

class OptionParser {
    std::string title_, examples_;
    std::vector<BaseOption*> opts_;
    BaseOption* find(const std::string& k) { for (auto* o : opts_) if (o->matches(k)) return o; return nullptr; }
    void apply(const std::string& key, const std::string* val) {
        BaseOption* o = find(key);
        if (!o) throw X_OptionError(key, "Option doesn't exist");
        if (o->has_arg()) {
            if (!val) throw X_OptionError(key, "Missing non-optional argument");
            if (!o->set_value(*val)) throw X_OptionError(key, "Couldn't set_value! valstr=\"" + *val + "\"");
        } else {
            o->set_value(val ? *val : std::string());
        }
    }
public:
    OptionParser(const std::string& t, const std::string& e) : title_(t), examples_(e) {}
    template <class T> void add(Option<T>& o) { opts_.push_back(&o); }
    void add(BaseOption& o) { opts_.push_back(&o); }
    void usage() const {
        std::cerr << "\n" << title_ << "\n\nUsage: " << examples_ << "\n";
        for (auto* o : opts_) std::cerr << "\t" << o->key() << "\t" << o->help_text() << "\n";
        std::cerr << std::endl;
    }
    bool check_compulsory_arguments(bool verbose = false) const {
        bool ok = true;
        for (auto* o : opts_)
            if (o->compulsory() && o->unset()) { ok = false; if (verbose) std::cerr << "***: " << o->key() << " is compulsory\n"; }
        return ok;
    }
    // returns the index of the first non-option argument
    unsigned int parse_command_line(unsigned int argc, char** argv, int skip = 0, bool = false) {
        unsigned int a = 1 + skip;
        while (a < argc) {
            std::string arg(argv[a]);
            if (arg.empty() || arg[0] != '-') break;
            std::string key = arg, val;
            bool hasval = false;
            std::size_t eq = arg.find('=');
            if (eq != std::string::npos) { key = arg.substr(0, eq); val = arg.substr(eq + 1); hasval = true; }
            BaseOption* o = find(key);
            if (!o) throw X_OptionError(key, "Option doesn't exist");
            if (o->has_arg() && !hasval) {
                if (a + 1 >= argc) throw X_OptionError(key, "Missing non-optional argument");
                val = argv[++a]; hasval = true;
            }
            apply(key, hasval ? &val : nullptr);
            ++a;
        }
        return a;
    }
    void parse_config_file(const std::string& filename) {
        std::ifstream f(filename.c_str());
        if (!f) throw X_OptionError(filename, "Couldn't open the file");
        std::string line;
        while (std::getline(f, line)) {
            std::size_t h = line.find('#');
            if (h != std::string::npos) line.erase(h);
            std::stringstream ss(line);
            std::string tok;
            while (ss >> tok) {
                std::string key = tok, val;
                bool hasval = false;
                std::size_t eq = tok.find('=');
                if (eq != std::string::npos) { key = tok.substr(0, eq); val = tok.substr(eq + 1); hasval = true; }
                BaseOption* o = find(key);
                if (!o) throw X_OptionError(key, "Option doesn't exist");
                if (o->has_arg() && !hasval) { if (!(ss >> val)) throw X_OptionError(key, "Missing non-optional argument"); hasval = true; }
                apply(key, hasval ? &val : nullptr);
            }
        }
    }
};

}  // namespace Utilities
#endif
