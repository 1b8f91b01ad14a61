// yaml.cpp — see yaml.hpp.
#include "yaml.hpp"

#include <fstream>
#include <sstream>
#include <stdexcept>

namespace sfe
{
namespace yaml
{
namespace
{
struct Line
{
    size_t indent;
    std::string text; // no leading spaces, no comment, right-trimmed
    size_t number;
};

[[noreturn]] void fail(const std::string &what, size_t line)
{
    throw std::invalid_argument("YAML: " + what + " (Line " + std::to_string(line) + ")");
}

std::string rtrim(std::string s)
{
    while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.pop_back();
    return s;
}

std::string trim(const std::string &s)
{
    size_t a = 0;
    while (a < s.size() && (s[a] == ' ' || s[a] == '\t')) ++a;
    return rtrim(s.substr(a));
}

// strip a trailing comment (# at line start or after whitespace, outside quotes)
std::string strip_comment(const std::string &s)
{
    char quote = 0;
    for (size_t i = 0; i < s.size(); ++i)
    {
        const char c = s[i];
        if (quote != 0)
        {
            if (c == quote)
            {
                if (quote == '\'' && i + 1 < s.size() && s[i + 1] == '\'') ++i;
                else quote = 0;
            }
            else if (quote == '"' && c == '\\') ++i;
        }
        else if (c == '\'' || c == '"')
        {
            // a quote only opens a quoted scalar at the start of a token
            if (i == 0 || s[i - 1] == ' ' || s[i - 1] == '[' || s[i - 1] == '{' || s[i - 1] == ',' || s[i - 1] == ':')
                quote = c;
        }
        else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t'))
            return s.substr(0, i);
    }
    return s;
}

std::string unquote(const std::string &s, size_t line)
{
    if (s.size() >= 2 && s.front() == '\'' && s.back() == '\'')
    {
        std::string out;
        for (size_t i = 1; i + 1 < s.size(); ++i)
        {
            if (s[i] == '\'' && i + 2 < s.size() && s[i + 1] == '\'') ++i;
            out += s[i];
        }
        return out;
    }
    if (s.size() >= 2 && s.front() == '"' && s.back() == '"')
    {
        std::string out;
        for (size_t i = 1; i + 1 < s.size(); ++i)
        {
            if (s[i] == '\\' && i + 2 < s.size())
            {
                ++i;
                switch (s[i])
                {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                default: out += s[i]; break;
                }
            }
            else out += s[i];
        }
        return out;
    }
    if (!s.empty() && (s.front() == '\'' || s.front() == '"')) fail("unterminated quoted scalar", line);
    return s;
}

// position of the ':' that separates a block-mapping key from its value, or npos
size_t key_colon(const std::string &t)
{
    char quote = 0;
    int depth = 0;
    for (size_t i = 0; i < t.size(); ++i)
    {
        const char c = t[i];
        if (quote != 0)
        {
            if (c == quote)
            {
                if (quote == '\'' && i + 1 < t.size() && t[i + 1] == '\'') ++i;
                else quote = 0;
            }
            else if (quote == '"' && c == '\\') ++i;
            continue;
        }
        if ((c == '\'' || c == '"') && i == 0)
        {
            quote = c;
            continue;
        }
        if (i == 0 && (c == '[' || c == '{')) return std::string::npos; // a flow collection, not a key
        if (c == '[' || c == '{') ++depth;
        else if (c == ']' || c == '}') --depth;
        else if (c == ':' && depth <= 0 && (i + 1 == t.size() || t[i + 1] == ' ' || t[i + 1] == '\t')) return i;
    }
    return std::string::npos;
}

class FlowParser
{
public:
    FlowParser(const std::string &s, size_t line) : s_(s), line_(line) {}
    Node parse_all()
    {
        Node n = value();
        skip();
        if (i_ != s_.size()) fail("unexpected text after flow collection: '" + s_.substr(i_) + "'", line_);
        return n;
    }

private:
    const std::string &s_;
    size_t i_{0};
    size_t line_;

    void skip()
    {
        while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\t')) ++i_;
    }
    bool at_colon() const
    {
        return i_ < s_.size() && s_[i_] == ':' &&
                (i_ + 1 == s_.size() || s_[i_ + 1] == ' ' || s_[i_ + 1] == ',' || s_[i_ + 1] == ']' || s_[i_ + 1] == '}' ||
                        s_[i_ + 1] == '[' || s_[i_ + 1] == '{');
    }
    std::string scalar_token()
    {
        skip();
        const size_t start = i_;
        if (i_ < s_.size() && (s_[i_] == '\'' || s_[i_] == '"'))
        {
            const char q = s_[i_++];
            while (i_ < s_.size())
            {
                if (s_[i_] == q)
                {
                    if (q == '\'' && i_ + 1 < s_.size() && s_[i_ + 1] == '\'') { i_ += 2; continue; }
                    ++i_;
                    break;
                }
                if (q == '"' && s_[i_] == '\\') ++i_;
                ++i_;
            }
            return unquote(s_.substr(start, i_ - start), line_);
        }
        while (i_ < s_.size() && s_[i_] != ',' && s_[i_] != ']' && s_[i_] != '}' && s_[i_] != '[' && s_[i_] != '{' && !at_colon()) ++i_;
        return trim(s_.substr(start, i_ - start));
    }
    Node value()
    {
        skip();
        Node n;
        n.line = line_;
        if (i_ >= s_.size()) return n;
        if (s_[i_] == '[')
        {
            ++i_;
            n.kind = Node::Seq;
            for (;;)
            {
                skip();
                if (i_ >= s_.size()) fail("unterminated flow sequence", line_);
                if (s_[i_] == ']') { ++i_; break; }
                Node item = value();
                skip();
                if (at_colon())
                {
                    // "key: value" inside a flow sequence = a single-pair mapping
                    if (!item.is_scalar()) fail("complex keys are not supported", line_);
                    ++i_;
                    Node pair;
                    pair.kind = Node::Map;
                    pair.line = line_;
                    pair.map.emplace_back(item.scalar, value());
                    item = std::move(pair);
                    skip();
                }
                n.seq.push_back(std::move(item));
                if (i_ < s_.size() && s_[i_] == ',') { ++i_; continue; }
                if (i_ < s_.size() && s_[i_] == ']') { ++i_; break; }
                fail("expected ',' or ']' in flow sequence", line_);
            }
            return n;
        }
        if (s_[i_] == '{')
        {
            ++i_;
            n.kind = Node::Map;
            for (;;)
            {
                skip();
                if (i_ >= s_.size()) fail("unterminated flow mapping", line_);
                if (s_[i_] == '}') { ++i_; break; }
                const std::string key = scalar_token();
                skip();
                Node val;
                val.line = line_;
                if (at_colon())
                {
                    ++i_;
                    val = value();
                }
                n.map.emplace_back(key, std::move(val));
                skip();
                if (i_ < s_.size() && s_[i_] == ',') { ++i_; continue; }
                if (i_ < s_.size() && s_[i_] == '}') { ++i_; break; }
                fail("expected ',' or '}' in flow mapping", line_);
            }
            return n;
        }
        n.kind = Node::Scalar;
        n.scalar = scalar_token();
        return n;
    }
};

class BlockParser
{
public:
    explicit BlockParser(const std::string &text)
    {
        std::istringstream in(text);
        std::string raw;
        size_t number = 0;
        while (std::getline(in, raw))
        {
            ++number;
            std::string body = rtrim(strip_comment(raw));
            size_t indent = 0;
            while (indent < body.size() && body[indent] == ' ') ++indent;
            if (indent < body.size() && body[indent] == '\t') fail("tabs are not allowed for indentation", number);
            body = body.substr(indent);
            if (body.empty() || body == "---" || body == "...") continue;
            lines_.push_back({indent, body, number});
        }
    }
    Node parse_document()
    {
        if (lines_.empty()) return Node{};
        Node n = parse_block(lines_[0].indent);
        if (pos_ != lines_.size()) fail("unexpected indentation", lines_[pos_].number);
        return n;
    }

private:
    std::vector<Line> lines_;
    size_t pos_{0};

    static bool is_seq_item(const std::string &t) { return t == "-" || (t.size() > 1 && t[0] == '-' && t[1] == ' '); }

    // value text that starts on the current line; a flow collection may continue on the next lines
    Node parse_inline(const std::string &rest, size_t line)
    {
        Node n;
        n.line = line;
        if (rest[0] == '[' || rest[0] == '{')
        {
            std::string buf = rest;
            // gather until brackets balance
            for (;;)
            {
                int depth = 0;
                char quote = 0;
                for (size_t i = 0; i < buf.size(); ++i)
                {
                    const char c = buf[i];
                    if (quote != 0)
                    {
                        if (c == quote) quote = 0;
                        else if (quote == '"' && c == '\\') ++i;
                    }
                    else if ((c == '\'' || c == '"') && (i == 0 || buf[i - 1] == ' ' || buf[i - 1] == '[' || buf[i - 1] == '{' || buf[i - 1] == ',' || buf[i - 1] == ':'))
                        quote = c;
                    else if (c == '[' || c == '{') ++depth;
                    else if (c == ']' || c == '}') --depth;
                }
                if (depth <= 0) break;
                if (pos_ >= lines_.size()) fail("unterminated flow collection", line);
                buf += ' ';
                buf += lines_[pos_].text;
                ++pos_;
            }
            FlowParser fp(buf, line);
            return fp.parse_all();
        }
        n.kind = Node::Scalar;
        n.scalar = unquote(rest, line);
        return n;
    }

    Node parse_block(size_t indent)
    {
        const Line &l = lines_[pos_];
        if (is_seq_item(l.text)) return parse_seq(indent);
        if (key_colon(l.text) != std::string::npos) return parse_map(indent);
        // a bare scalar / flow collection as a whole block
        const std::string text = l.text;
        const size_t number = l.number;
        ++pos_;
        return parse_inline(text, number);
    }

    Node parse_map(size_t indent)
    {
        Node n;
        n.kind = Node::Map;
        n.line = lines_[pos_].number;
        while (pos_ < lines_.size() && lines_[pos_].indent == indent && !is_seq_item(lines_[pos_].text))
        {
            const Line l = lines_[pos_];
            const size_t colon = key_colon(l.text);
            if (colon == std::string::npos) fail("expected 'key: value'", l.number);
            const std::string key = unquote(trim(l.text.substr(0, colon)), l.number);
            const std::string rest = trim(l.text.substr(colon + 1));
            ++pos_;
            Node val;
            val.line = l.number;
            if (!rest.empty()) val = parse_inline(rest, l.number);
            else if (pos_ < lines_.size())
            {
                if (lines_[pos_].indent > indent) val = parse_block(lines_[pos_].indent);
                else if (lines_[pos_].indent == indent && is_seq_item(lines_[pos_].text)) val = parse_seq(indent);
            }
            n.map.emplace_back(key, std::move(val));
        }
        if (pos_ < lines_.size() && lines_[pos_].indent > indent) fail("unexpected indentation", lines_[pos_].number);
        return n;
    }

    Node parse_seq(size_t indent)
    {
        Node n;
        n.kind = Node::Seq;
        n.line = lines_[pos_].number;
        while (pos_ < lines_.size() && lines_[pos_].indent == indent && is_seq_item(lines_[pos_].text))
        {
            const Line l = lines_[pos_];
            std::string content = l.text.size() > 1 ? l.text.substr(2) : std::string();
            size_t extra = 2;
            while (!content.empty() && content[0] == ' ')
            {
                content.erase(0, 1);
                ++extra;
            }
            if (content.empty())
            {
                ++pos_;
                Node item;
                item.line = l.number;
                if (pos_ < lines_.size() && lines_[pos_].indent > indent) item = parse_block(lines_[pos_].indent);
                n.seq.push_back(std::move(item));
                continue;
            }
            if (is_seq_item(content) || (key_colon(content) != std::string::npos && content[0] != '[' && content[0] != '{'))
            {
                // compact nested collection: treat the rest of the line as a line of its own
                lines_[pos_] = {indent + extra, content, l.number};
                n.seq.push_back(parse_block(indent + extra));
                continue;
            }
            ++pos_;
            n.seq.push_back(parse_inline(content, l.number));
        }
        return n;
    }
};
} // namespace

Node parse(const std::string &text)
{
    BlockParser p(text);
    return p.parse_document();
}

Node parse_file(const std::string &path)
{
    std::ifstream in(path);
    if (!in) throw std::invalid_argument("failed to open " + path);
    std::ostringstream ss;
    ss << in.rdbuf();
    return parse(ss.str());
}
} // namespace yaml
} // namespace sfe
