// json.hpp — small ordered JSON value + parser for the flat JSON-lines front-end.
#ifndef SFE_JSON_HPP_
#define SFE_JSON_HPP_

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

namespace sfe
{
struct Json;
using JsonList = std::vector<Json>;
using JsonMap = std::vector<std::pair<std::string, Json>>;

struct Json
{
    std::variant<std::monostate, bool, long long, double, std::string, std::shared_ptr<JsonList>,
            std::shared_ptr<JsonMap>>
            v;
    bool is_bool() const { return std::holds_alternative<bool>(v); }
    bool is_int() const { return std::holds_alternative<long long>(v); }
    bool is_double() const { return std::holds_alternative<double>(v); }
    bool is_string() const { return std::holds_alternative<std::string>(v); }
    bool is_list() const { return std::holds_alternative<std::shared_ptr<JsonList>>(v); }
    bool is_map() const { return std::holds_alternative<std::shared_ptr<JsonMap>>(v); }
    bool boolean() const { return std::get<bool>(v); }
    long long integer() const { return std::get<long long>(v); }
    double number() const { return is_int() ? static_cast<double>(integer()) : std::get<double>(v); }
    const std::string &str() const { return std::get<std::string>(v); }
    const JsonList &list() const { return *std::get<std::shared_ptr<JsonList>>(v); }
    const JsonMap &map() const { return *std::get<std::shared_ptr<JsonMap>>(v); }
    const Json *find(const char *key) const
    {
        if (!is_map()) return nullptr;
        for (const auto &kv : map())
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const Json &at(const char *key) const
    {
        const Json *p = find(key);
        if (p == nullptr) throw std::runtime_error(std::string("flat description: missing key '") + key + "'");
        return *p;
    }
};

class JsonReader
{
public:
    JsonReader(const char *begin, const char *end) : p_(begin), end_(end) {}
    Json parse()
    {
        skip();
        if (p_ >= end_) fail("unexpected end");
        Json out;
        const char c = *p_;
        if (c == '{')
        {
            ++p_;
            auto m = std::make_shared<JsonMap>();
            skip();
            if (peek() == '}') { ++p_; out.v = m; return out; }
            for (;;)
            {
                Json key = parse();
                skip();
                expect(':');
                m->emplace_back(key.str(), parse());
                skip();
                if (peek() == ',') { ++p_; continue; }
                expect('}');
                break;
            }
            out.v = m;
        }
        else if (c == '[')
        {
            ++p_;
            auto l = std::make_shared<JsonList>();
            skip();
            if (peek() == ']') { ++p_; out.v = l; return out; }
            for (;;)
            {
                l->push_back(parse());
                skip();
                if (peek() == ',') { ++p_; continue; }
                expect(']');
                break;
            }
            out.v = l;
        }
        else if (c == '"')
        {
            ++p_;
            std::string s;
            while (p_ < end_ && *p_ != '"')
            {
                if (*p_ == '\\' && p_ + 1 < end_)
                {
                    ++p_;
                    if (*p_ == 'n') s += '\n';
                    else if (*p_ == 't') s += '\t';
                    else if (*p_ == 'u' && p_ + 4 < end_)
                    {
                        s += static_cast<char>(std::strtoul(std::string(p_ + 1, p_ + 5).c_str(), nullptr, 16) & 0x7f);
                        p_ += 4;
                    }
                    else s += *p_;
                    ++p_;
                }
                else s += *p_++;
            }
            expect('"');
            out.v = std::move(s);
        }
        else if (match("true")) out.v = true;
        else if (match("false")) out.v = false;
        else if (match("null")) {}
        else
        {
            const char *q = p_;
            bool real = false;
            while (q < end_ && (std::isalnum(static_cast<unsigned char>(*q)) || *q == '-' || *q == '+' || *q == '.'))
            {
                if (*q == '.' || std::isalpha(static_cast<unsigned char>(*q))) real = true;
                ++q;
            }
            if (q == p_) fail("unexpected character");
            const std::string tok(p_, q);
            p_ = q;
            if (real) out.v = std::strtod(tok.c_str(), nullptr);
            else out.v = std::strtoll(tok.c_str(), nullptr, 10);
        }
        return out;
    }

private:
    const char *p_;
    const char *end_;
    char peek() const { return p_ < end_ ? *p_ : '\0'; }
    void skip()
    {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_;
    }
    void expect(char c)
    {
        if (peek() != c) fail(std::string("expected '") + c + "'");
        ++p_;
    }
    bool match(const char *word)
    {
        const size_t n = std::strlen(word);
        if (static_cast<size_t>(end_ - p_) >= n && std::strncmp(p_, word, n) == 0) { p_ += n; return true; }
        return false;
    }
    [[noreturn]] void fail(const std::string &why) const { throw std::runtime_error("flat description: JSON " + why); }
};
} // namespace sfe
#endif
