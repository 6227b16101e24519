// C wrapper around the REFERENCE's own pose-graph code.  graph.h / graph.cpp are
// compiled where they lie under /root/reference (never copied); this file only
// adapts them to ctypes.  Output goes to oracle/_ref/libgraph_ref.so.
// Test infrastructure only.
#include <cstdint>
#include <vector>

#include "graph.h"  // -I/root/reference/map_merge_3d/src

extern "C" int ref_graph(int n, const int32_t* st, const double* conf, double thr, int32_t* in_component,
                         int32_t* tree_edges, int* n_tree_edges, int32_t* centers, int* n_centers, int* n_nodes)
{
  std::vector<TransformEstimate> est;
  for (int i = 0; i < n; ++i) {
    TransformEstimate e((size_t)st[2 * i], (size_t)st[2 * i + 1]);
    e.confidence = conf[i];
    est.push_back(e);
  }
  *n_nodes = (int)numberOfNodesInEstimates(est);
  std::vector<TransformEstimate> comp = largestConnectedComponent(est, thr);
  size_t c = 0;
  for (int i = 0; i < n; ++i) {
    in_component[i] = 0;
    if (c < comp.size() && comp[c].source_idx == est[i].source_idx && comp[c].target_idx == est[i].target_idx) {
      in_component[i] = 1;
      ++c;
    }
  }
  Graph tree;
  std::vector<size_t> cen;
  *n_tree_edges = 0;
  *n_centers = 0;
  if (comp.empty()) return 0;
  findMaxSpanningTree(comp, tree, cen);
  int ne = 0;
  tree.forEach([&](const GraphEdge& e) {
    tree_edges[2 * ne] = (int)e.from;
    tree_edges[2 * ne + 1] = (int)e.to;
    ++ne;
  });
  *n_tree_edges = ne;
  for (size_t i = 0; i < cen.size(); ++i) centers[i] = (int)cen[i];
  *n_centers = (int)cen.size();
  return 0;
}
