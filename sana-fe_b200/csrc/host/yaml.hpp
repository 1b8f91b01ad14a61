// yaml.hpp — in-tree reader for the YAML subset the Architecture / SNN description
// files use (the reference links RapidYAML v0.13.0, fetched from the network by its
// CMake and unavailable here). Supported: block mappings and sequences (including
// sequences at the indentation of their parent key and "- key: value" compact
// entries), flow sequences / mappings spanning several lines, single-pair mappings
// inside flow sequences ("[type: dense, weight: [1, 2]]"), single- and double-quoted
// scalars, plain scalars, comments. Not supported (unused by the formats): anchors,
// tags, block scalars, multi-document streams, complex keys.
#ifndef SFE_YAML_HPP_
#define SFE_YAML_HPP_

#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace sfe
{
namespace yaml
{
struct Node
{
    enum Kind
    {
        Null,
        Scalar,
        Map,
        Seq
    };
    Kind kind{Null};
    std::string scalar;                             // Scalar
    std::vector<std::pair<std::string, Node>> map;  // Map, in document order
    std::vector<Node> seq;                          // Seq
    size_t line{0};                                 // 1-based source line (error messages)

    bool is_map() const { return kind == Map; }
    bool is_seq() const { return kind == Seq; }
    bool is_scalar() const { return kind == Scalar; }
    bool is_null() const { return kind == Null; }
    const Node *find(const std::string &key) const
    {
        if (kind != Map) return nullptr;
        for (const auto &kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

// Parses a whole document; throws std::invalid_argument with a line number on error.
Node parse(const std::string &text);
Node parse_file(const std::string &path);
} // namespace yaml
} // namespace sfe
#endif
