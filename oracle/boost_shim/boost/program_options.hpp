// shim of <boost/program_options.hpp>: only what /root/reference/src/source.cc uses (:157-185, :1116-1176).
// TEST INFRASTRUCTURE ONLY, written from scratch from Boost's documented behaviour:
//   * options_description / add_options()(name, value_semantic*, help)
//   * value<T>(T*) with ->multitoken() / ->default_value(v); bool_switch(bool*)
//   * command_line_parser(argc, argv).extra_style_parser(f).options(desc).run()
//   * store(parsed, vm); notify(vm); vm.count(name); operator<<(ostream, options_description)
// Command-line semantics follow boost::program_options::detail::cmdline::run():
//   style parsers are tried in order (the extra parser first, then long options "--name[=v]",
//   then short "-x"); a key option first takes its min_tokens from the following tokens, a second
//   sweep hands further nameless (positional) tokens to options that can take more; nameless
//   options that remain are skipped by store() because no positional description is given.
#pragma once
#include <climits>
#include <cstring>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "lexical_cast.hpp"

namespace boost {
namespace program_options {

struct option {
    std::string string_key;
    int position_key;
    std::vector<std::string> value;
    std::vector<std::string> original_tokens;
    bool unregistered, case_insensitive;
    option() : position_key(-1), unregistered(false), case_insensitive(false) {}
};

class error : public std::logic_error {
public:
    explicit error(const std::string &m) : std::logic_error(m) {}
};

class value_semantic {
public:
    virtual ~value_semantic() {}
    virtual unsigned min_tokens() const = 0;
    virtual unsigned max_tokens() const = 0;
    virtual bool is_composing() const = 0;                       // vector-valued: occurrences append
    virtual void parse(const std::vector<std::string> &tokens, const std::string &name, bool first) = 0;
    virtual void apply_default() = 0;
    virtual std::string default_text() const = 0;
};

namespace shim_detail {
template <class T> struct elem {
    typedef T type;
    static const bool is_vector = false;
    static void assign(T &dst, const std::vector<T> &v, bool) { dst = v.back(); }
};
template <class T> struct elem<std::vector<T> > {
    typedef T type;
    static const bool is_vector = true;
    static void assign(std::vector<T> &dst, const std::vector<T> &v, bool first)
    {
        if (first) dst.clear();
        dst.insert(dst.end(), v.begin(), v.end());
    }
};
}  // namespace shim_detail

template <class T> class typed_value : public value_semantic {
    typedef typename shim_detail::elem<T>::type E;
public:
    explicit typed_value(T *store_to) : store_(store_to), multitoken_(false), zero_tokens_(false), has_default_(false) {}
    typed_value *multitoken() { multitoken_ = true; return this; }
    typed_value *zero_tokens() { zero_tokens_ = true; return this; }
    typed_value *default_value(const T &v)
    {
        default_ = v; has_default_ = true;
        std::ostringstream os; print(os, v); default_text_ = os.str();
        return this;
    }
    unsigned min_tokens() const override { return zero_tokens_ ? 0 : 1; }
    unsigned max_tokens() const override { return multitoken_ ? 32000 : (zero_tokens_ ? 0 : 1); }
    bool is_composing() const override { return shim_detail::elem<T>::is_vector; }
    void parse(const std::vector<std::string> &tokens, const std::string &name, bool first) override
    {
        if (zero_tokens_) { set_true(); return; }
        std::vector<E> vals;
        for (size_t i = 0; i < tokens.size(); i++) {
            try { vals.push_back(boost::lexical_cast<E>(tokens[i])); }
            catch (const bad_lexical_cast &) {
                throw error("the argument ('" + tokens[i] + "') for option '--" + name + "' is invalid");
            }
        }
        if (vals.empty()) throw error("the required argument for option '--" + name + "' is missing");
        if (store_) shim_detail::elem<T>::assign(*store_, vals, first);
    }
    void apply_default() override { if (has_default_ && store_) *store_ = default_; }
    std::string default_text() const override { return has_default_ ? default_text_ : std::string(); }
private:
    template <class U> static void print(std::ostream &os, const U &v) { os << v; }
    template <class U> static void print(std::ostream &os, const std::vector<U> &v) { for (size_t i = 0; i < v.size(); i++) os << (i ? " " : "") << v[i]; }
    void set_true() { set_true_impl(store_); }
    static void set_true_impl(bool *p) { if (p) *p = true; }
    template <class U> static void set_true_impl(U *) {}
    T *store_;
    bool multitoken_, zero_tokens_, has_default_;
    T default_;
    std::string default_text_;
};

template <class T> typed_value<T> *value(T *v) { return new typed_value<T>(v); }
template <class T> typed_value<T> *value() { return new typed_value<T>(nullptr); }
inline typed_value<bool> *bool_switch(bool *v) { typed_value<bool> *t = new typed_value<bool>(v); t->zero_tokens(); return t; }

struct option_description {
    std::string long_name, short_name, description;
    std::shared_ptr<value_semantic> semantic;
};

class options_description;
class options_description_easy_init {
public:
    explicit options_description_easy_init(options_description *o) : owner_(o) {}
    options_description_easy_init &operator()(const char *name, const char *description);
    options_description_easy_init &operator()(const char *name, value_semantic *s, const char *description);
private:
    options_description *owner_;
};

class options_description {
public:
    explicit options_description(const std::string &caption = std::string()) : caption_(caption) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    void add(const char *name, value_semantic *s, const char *description)
    {
        option_description d;
        std::string n(name);
        const size_t comma = n.find(',');
        d.long_name = n.substr(0, comma);
        if (comma != std::string::npos) d.short_name = n.substr(comma + 1);
        d.description = description;
        d.semantic.reset(s);
        options_.push_back(d);
    }
    // exact long name, else unambiguous prefix ("guessing" is on in the default style); nullptr if unknown
    const option_description *find_nothrow(const std::string &name, bool approx = true) const
    {
        for (size_t i = 0; i < options_.size(); i++) if (options_[i].long_name == name) return &options_[i];
        if (!approx) return nullptr;
        std::vector<const option_description *> hits;
        for (size_t i = 0; i < options_.size(); i++)
            if (options_[i].long_name.compare(0, name.size(), name) == 0) hits.push_back(&options_[i]);
        if (hits.size() == 1) return hits[0];
        if (hits.size() > 1) {
            std::string m = "option '--" + name + "' is ambiguous and matches ";
            for (size_t i = 0; i < hits.size(); i++) {
                if (i) m += (i + 1 == hits.size()) ? ", and " : ", ";
                m += "'--" + hits[i]->long_name + "'";
            }
            throw error(m);
        }
        return nullptr;
    }
    const option_description *find_short(const std::string &s) const
    {
        for (size_t i = 0; i < options_.size(); i++) if (!s.empty() && options_[i].short_name == s) return &options_[i];
        return nullptr;
    }
    const std::vector<option_description> &options() const { return options_; }
    const std::string &caption() const { return caption_; }
private:
    std::string caption_;
    std::vector<option_description> options_;
};

inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *description)
{
    typed_value<bool> *t = new typed_value<bool>(nullptr);
    t->zero_tokens();
    owner_->add(name, t, description);
    return *this;
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, value_semantic *s,
                                                                                  const char *description)
{
    owner_->add(name, s, description);
    return *this;
}

inline std::ostream &operator<<(std::ostream &os, const options_description &d)
{
    if (!d.caption().empty()) os << d.caption() << ":\n";
    for (size_t i = 0; i < d.options().size(); i++) {
        const option_description &o = d.options()[i];
        std::string head = "  ";
        if (!o.short_name.empty()) head += "-" + o.short_name + " [ --" + o.long_name + " ]";
        else head += "--" + o.long_name;
        if (o.semantic->max_tokens() > 0) {
            head += " arg";
            if (!o.semantic->default_text().empty()) head += " (=" + o.semantic->default_text() + ")";
        }
        if (head.size() < 28) head.resize(28, ' ');
        os << head << o.description << "\n";
    }
    return os;
}

struct parsed_options {
    std::vector<option> options;
    const options_description *description;
    parsed_options() : description(nullptr) {}
};

class variables_map {
public:
    size_t count(const std::string &name) const { return seen_.count(name); }
    std::map<std::string, int> seen_;
};

typedef std::vector<option> (*ext_parser)(std::vector<std::string> &);

class command_line_parser {
public:
    command_line_parser(int argc, char *argv[]) : desc_(nullptr), extra_(nullptr)
    {
        for (int i = 1; i < argc; i++) args_.push_back(argv[i]);
    }
    command_line_parser &extra_style_parser(ext_parser p) { extra_ = p; return *this; }
    command_line_parser &options(const options_description &d) { desc_ = &d; return *this; }

    parsed_options run()
    {
        std::vector<std::string> args = args_;
        std::vector<option> result;
        while (!args.empty()) {
            bool ok = false;
            for (int sp = 0; sp < 3 && !ok; sp++) {
                const size_t before = args.size();
                std::vector<option> next = style_parse(sp, args);
                if (!next.empty()) {
                    std::vector<std::string> none;
                    for (size_t k = 0; k + 1 < next.size(); k++) finish_option(next[k], none);
                    finish_option(next.back(), args);
                    result.insert(result.end(), next.begin(), next.end());
                }
                if (args.size() != before) ok = true;
            }
            if (!ok) {                                   // a plain token: nameless (positional) option
                option opt;
                opt.value.push_back(args[0]);
                opt.original_tokens.push_back(args[0]);
                result.push_back(opt);
                args.erase(args.begin());
            }
        }
        // key options that can take more tokens absorb the nameless options that follow them
        std::vector<option> result2;
        for (size_t i = 0; i < result.size(); i++) {
            result2.push_back(result[i]);
            option &opt = result2.back();
            if (opt.string_key.empty()) continue;
            const option_description *xd = desc_->find_nothrow(opt.string_key, false);
            if (!xd) continue;
            const unsigned mn = xd->semantic->min_tokens(), mx = xd->semantic->max_tokens();
            if (mn < mx && opt.value.size() < mx) {
                int can_take = (int)mx - (int)opt.value.size();
                size_t j = i + 1;
                for (; can_take && j < result.size(); --can_take, ++j) {
                    option &o2 = result[j];
                    if (!o2.string_key.empty() || o2.position_key == INT_MAX) break;
                    opt.value.push_back(o2.value[0]);
                    opt.original_tokens.push_back(o2.original_tokens[0]);
                }
                i = j - 1;
            }
        }
        int pos = 0;
        for (size_t i = 0; i < result2.size(); i++) if (result2[i].string_key.empty()) result2[i].position_key = pos++;
        parsed_options out;
        out.options.swap(result2);
        out.description = desc_;
        return out;
    }

private:
    std::vector<option> style_parse(int which, std::vector<std::string> &args)
    {
        std::vector<option> r;
        if (which == 0) { if (extra_) r = extra_(args); return r; }
        const std::string &tok = args[0];
        if (which == 1 && tok.size() >= 3 && tok[0] == '-' && tok[1] == '-') {          // --name[=value]
            option o;
            const size_t eq = tok.find('=');
            o.string_key = tok.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos) o.value.push_back(tok.substr(eq + 1));
            o.original_tokens.push_back(tok);
            r.push_back(o);
            args.erase(args.begin());
        } else if (which == 2 && tok.size() >= 2 && tok[0] == '-' && tok[1] != '-') {    // -x[value]
            const option_description *d = desc_->find_short(tok.substr(1, 1));
            if (!d) throw error("unrecognised option '" + tok + "'");
            option o;
            o.string_key = d->long_name;
            if (tok.size() > 2) o.value.push_back(tok.substr(2));
            o.original_tokens.push_back(tok);
            r.push_back(o);
            args.erase(args.begin());
        }
        return r;
    }

    void finish_option(option &opt, std::vector<std::string> &other)
    {
        if (opt.string_key.empty()) return;
        const option_description *xd = desc_->find_nothrow(opt.string_key, true);
        if (!xd) throw error("unrecognised option '--" + opt.string_key + "'");
        opt.string_key = xd->long_name;
        unsigned mn = xd->semantic->min_tokens();
        const unsigned mx = xd->semantic->max_tokens();
        const size_t present = opt.value.size() + other.size();
        if (present < mn) throw error("the required argument for option '--" + opt.string_key + "' is missing");
        if (!opt.value.empty() && mx == 0) throw error("option '--" + opt.string_key + "' does not take any arguments");
        mn = opt.value.size() <= mn ? mn - (unsigned)opt.value.size() : 0;
        for (; !other.empty() && mn--;) {
            // a following token that is itself a known key option is not a value
            const std::string &t = other[0];
            if (t.size() >= 3 && t[0] == '-' && t[1] == '-' && desc_->find_nothrow(t.substr(2, t.find('=') - 2), false))
                throw error("the required argument for option '--" + opt.string_key + "' is missing");
            opt.value.push_back(t);
            opt.original_tokens.push_back(t);
            other.erase(other.begin());
        }
    }

    std::vector<std::string> args_;
    const options_description *desc_;
    ext_parser extra_;
};

inline void store(const parsed_options &p, variables_map &vm)
{
    for (size_t i = 0; i < p.options.size(); i++) {
        const option &o = p.options[i];
        if (o.string_key.empty()) continue;                      // nameless positional: skipped
        const option_description *d = p.description->find_nothrow(o.string_key, false);
        if (!d) throw error("unrecognised option '--" + o.string_key + "'");
        const bool first = vm.seen_.count(d->long_name) == 0;
        if (!first && !d->semantic->is_composing())
            throw error("option '--" + d->long_name + "' cannot be specified more than once");
        d->semantic->parse(o.value, d->long_name, first);
        vm.seen_[d->long_name]++;
    }
    for (size_t i = 0; i < p.description->options().size(); i++) {
        const option_description &d = p.description->options()[i];
        if (!vm.seen_.count(d.long_name) && !d.semantic->default_text().empty()) {
            d.semantic->apply_default();
            vm.seen_[d.long_name]++;          // defaulted options count() as 1, like Boost
        }
    }
}

inline void notify(variables_map &) {}

}  // namespace program_options
}  // namespace boost
