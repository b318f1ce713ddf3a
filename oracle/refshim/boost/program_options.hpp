// Stand-in for <boost/program_options.hpp> (see ../README.md): the subset src/kmers-build-signatures.cc:36-69
// uses — long options "--name value" (and the short letter after the comma), vector-valued options that may be
// repeated or, with multitoken(), followed by several values; "--help".
#pragma once
#include <cstdlib>
#include <filesystem>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace boost { namespace program_options {

struct value_semantic {
    virtual ~value_semantic() = default;
    virtual void add(const std::string &text) = 0;
    bool multi = false;
    value_semantic *multitoken() { multi = true; return this; }
};
template <class T> struct typed_value : value_semantic {
    T *dst;
    explicit typed_value(T *d) : dst(d) {}
    void add(const std::string &text) override { std::istringstream in(text); in >> *dst; }
};
template <> inline void typed_value<std::string>::add(const std::string &text) { *dst = text; }
template <> inline void typed_value<std::filesystem::path>::add(const std::string &text) { *dst = text; }
template <class T> struct typed_value<std::vector<T>> : value_semantic {
    std::vector<T> *dst;
    explicit typed_value(std::vector<T> *d) : dst(d) {}
    void add(const std::string &text) override { dst->emplace_back(text); }
};
template <class T> typed_value<T> *value(T *dst) { return new typed_value<T>(dst); }

struct option_entry { std::string long_name, short_name, description; std::shared_ptr<value_semantic> sem; };

class options_description {
public:
    explicit options_description(const std::string &caption) : caption_(caption) {}
    struct adder {
        options_description *d;
        adder &operator()(const char *name, value_semantic *sem, const char *desc) { d->add(name, sem, desc); return *this; }
        adder &operator()(const char *name, const char *desc) { d->add(name, nullptr, desc); return *this; }
    };
    adder add_options() { return adder{this}; }
    void add(const char *name, value_semantic *sem, const char *desc) {
        std::string n = name, s;
        const auto comma = n.find(',');
        if (comma != std::string::npos) { s = n.substr(comma + 1); n = n.substr(0, comma); }
        entries.push_back(option_entry{n, s, desc, std::shared_ptr<value_semantic>(sem)});
    }
    const option_entry *find(const std::string &arg) const {
        for (const auto &e : entries)
            if (arg == "--" + e.long_name || (!e.short_name.empty() && arg == "-" + e.short_name)) return &e;
        return nullptr;
    }
    std::string caption_;
    std::vector<option_entry> entries;
};
inline std::ostream &operator<<(std::ostream &os, const options_description &d) {
    os << d.caption_ << ":\n";
    for (const auto &e : d.entries) os << "  --" << e.long_name << "  " << e.description << "\n";
    return os;
}

class variables_map {
public:
    std::size_t count(const std::string &name) const { auto it = seen.find(name); return it == seen.end() ? 0 : it->second; }
    std::map<std::string, std::size_t> seen;
};
struct parsed_options { std::vector<std::pair<const option_entry *, std::string>> values; std::vector<const option_entry *> flags; };

inline parsed_options parse_command_line(int argc, char **argv, const options_description &desc) {
    parsed_options out;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        const option_entry *e = desc.find(a);
        if (!e) throw std::runtime_error("unrecognised option '" + a + "'");
        if (!e->sem) { out.flags.push_back(e); continue; }
        if (i + 1 >= argc) throw std::runtime_error("the required argument for option '" + a + "' is missing");
        out.values.emplace_back(e, argv[++i]);
        while (e->sem->multi && i + 1 < argc && argv[i + 1][0] != '-') out.values.emplace_back(e, argv[++i]);
    }
    return out;
}
inline void store(const parsed_options &p, variables_map &vm) {
    for (const auto &v : p.values) { v.first->sem->add(v.second); ++vm.seen[v.first->long_name]; }
    for (const auto *f : p.flags) ++vm.seen[f->long_name];
}
inline void notify(variables_map &) {}

}}  // namespace boost::program_options
