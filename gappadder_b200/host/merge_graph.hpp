// merge_graph.hpp -- the overlap graph and path search behind ContigsMerger's merged contigs
// (ContigsCompactor.cpp:724-983,1422-1454; GraphUtils.cpp:625-859,1028-1344), restated on node
// indices.  The reference orders several containers by POINTER value (std::set<AbstractGraphNode*>,
// std::set<std::vector<AbstractGraphNode*>>); its nodes are allocated in creation order, so index
// order stands in for pointer order here ("hard parts").
#pragma once
#include <string>
#include <vector>

namespace gpm {

class OverlapGraph {
public:
    struct Edge { int to; double len; };

    explicit OverlapGraph(int n_nodes) : adj_(n_nodes) {}
    int size() const { return (int)adj_.size(); }
    // GraphNodeRefExt::AddNgbrRef: edges keep insertion order (it drives the DFS and the path DP)
    void add_edge(int from, int to, double len) { adj_[from].push_back(Edge{to, len}); }
    const std::vector<Edge>& edges(int v) const { return adj_[v]; }
    bool is_neighbor(int from, int to) const;

    // AbstractGraph::SCC (GraphUtils.cpp:1028-1178): Tarjan from node 0..n-1, components returned in
    // REVERSE order of completion, each as an ascending list (std::set of pointers).
    std::vector<std::vector<int>> scc() const;

    // AbstractGraph::FindSimplePathsTopSort (GraphUtils.cpp:625-771) -> the sorted, duplicate-free set
    // of paths (lexicographic by node index, as std::set<vector<ptr>>).
    std::vector<std::vector<int>> find_paths(int max_per_root) const;

    // AbstractGraph::OutputGML (GraphUtils.cpp:1196-1253), byte for byte.
    std::string gml(const std::vector<std::string>& names) const;

private:
    std::vector<int> terminals(bool start, const std::vector<std::vector<int>>& sccs) const;
    std::vector<std::vector<int>> paths_from(int root, const std::vector<int>& order,
                                             const std::vector<int>& rank, const std::vector<int>& ends) const;
    std::vector<std::vector<Edge>> adj_;
};

// ContigsCompactor::RemoveDupRevCompPaths (ContigsCompactor.cpp:1422-1454): a path is dropped when the
// reverse-complement image of it (reversed, every node replaced by its partner node) occurs EARLIER
// in the sorted set.  rc_partner[v] = v ^ 1 for the [c0, c0_R, c1, c1_R, ...] node layout.
std::vector<std::vector<int>> remove_revcomp_duplicates(const std::vector<std::vector<int>>& paths);

} // namespace gpm
