// graph.cpp — host-side pose graph: largest connected component, maximum
// spanning tree, tree centre and chaining of pairwise transforms.
//
// Mirrors the behaviour (including the quirks SURVEY.md Appendix B lists) of
//   map_merge_3d/src/graph.cpp:64-102   largestConnectedComponent
//   map_merge_3d/src/graph.cpp:104-175  findMaxSpanningTree
//   map_merge_3d/src/map_merging.cpp:137-186  getTransform / computeGlobalTransforms
// At most a few hundred edges: this stays on the host; it must agree exactly with
// the reference because the tree decides every global transform.
#include <algorithm>
#include <deque>
#include <functional>

#include "mm3d_internal.cuh"

namespace mm3d {

namespace {

struct UnionFind {
  std::vector<size_t> parent, count, rank;
  explicit UnionFind(size_t n) : parent(n), count(n, 1), rank(n, 0)
  {
    for (size_t i = 0; i < n; ++i) parent[i] = i;
  }
  size_t root(size_t e)
  {
    size_t r = e;
    while (parent[r] != r) r = parent[r];
    while (parent[e] != e) {  // path compression
      const size_t nx = parent[e];
      parent[e] = r;
      e = nx;
    }
    return r;
  }
  // union by rank; on equal rank the SECOND set becomes the root (graph.cpp:47-61)
  void join(size_t a, size_t b)
  {
    if (rank[a] < rank[b]) { parent[a] = b; count[b] += count[a]; }
    else if (rank[b] < rank[a]) { parent[b] = a; count[a] += count[b]; }
    else { parent[a] = b; rank[b]++; count[b] += count[a]; }
  }
};

struct WEdge {
  size_t from, to;
  double weight;
  bool operator>(const WEdge& o) const { return weight > o.weight; }
};

size_t node_count(const std::vector<HostEstimate>& e)
{
  size_t n = 0;
  for (const HostEstimate& x : e) n = std::max(n, std::max(x.source_idx + 1, x.target_idx + 1));
  return n;
}

template <typename F>
void breadth_first(const std::vector<std::vector<WEdge>>& adj, size_t start, F visit)
{
  std::vector<char> seen(adj.size(), 0);
  std::deque<size_t> q;
  seen[start] = 1;
  q.push_back(start);
  while (!q.empty()) {
    const size_t v = q.front();
    q.pop_front();
    for (const WEdge& e : adj[v]) {
      if (seen[e.to]) continue;
      visit(e);
      seen[e.to] = 1;
      q.push_back(e.to);
    }
  }
}

void mat_mul(const float* a, const float* b, float* r)
{
  for (int i = 0; i < 4; ++i)
    for (int c = 0; c < 4; ++c) {
      float acc = a[i * 4 + 0] * b[0 * 4 + c];
      acc += a[i * 4 + 1] * b[1 * 4 + c];
      acc += a[i * 4 + 2] * b[2 * 4 + c];
      acc += a[i * 4 + 3] * b[3 * 4 + c];
      r[i * 4 + c] = acc;
    }
}

// general 4x4 inverse (Eigen::Matrix4f::inverse()): adjugate / determinant
void mat_inverse(const float* m, float* out)
{
  float c[16];
  auto minor3 = [&](int r0, int r1, int r2, int c0, int c1, int c2) {
    return m[r0 * 4 + c0] * (m[r1 * 4 + c1] * m[r2 * 4 + c2] - m[r1 * 4 + c2] * m[r2 * 4 + c1]) -
           m[r0 * 4 + c1] * (m[r1 * 4 + c0] * m[r2 * 4 + c2] - m[r1 * 4 + c2] * m[r2 * 4 + c0]) +
           m[r0 * 4 + c2] * (m[r1 * 4 + c0] * m[r2 * 4 + c1] - m[r1 * 4 + c1] * m[r2 * 4 + c0]);
  };
  for (int r = 0; r < 4; ++r)
    for (int cc = 0; cc < 4; ++cc) {
      int rr[3], cl[3];
      for (int k = 0, t = 0; k < 4; ++k)
        if (k != r) rr[t++] = k;
      for (int k = 0, t = 0; k < 4; ++k)
        if (k != cc) cl[t++] = k;
      const float mn = minor3(rr[0], rr[1], rr[2], cl[0], cl[1], cl[2]);
      c[cc * 4 + r] = ((r + cc) & 1) ? -mn : mn;  // transposed cofactor
    }
  const float det = m[0] * c[0] + m[1] * c[4] + m[2] * c[8] + m[3] * c[12];
  for (int i = 0; i < 16; ++i) out[i] = c[i] / det;
}

}  // namespace

std::vector<std::vector<float>> compute_global_transforms(const std::vector<HostEstimate>& pairwise, double confidence_threshold,
                                                          int* reference_frame, std::vector<int>* in_component,
                                                          std::vector<std::pair<int, int>>* tree_edges, std::vector<int>* centers_out)
{
  const size_t nodes = node_count(pairwise);
  std::vector<std::vector<float>> global(nodes, std::vector<float>(16, 0.0f));
  if (reference_frame) *reference_frame = -1;
  if (in_component) in_component->assign(pairwise.size(), 0);
  if (tree_edges) tree_edges->clear();
  if (centers_out) centers_out->clear();
  if (nodes == 0) return global;

  // largest connected component over edges with confidence >= threshold
  UnionFind comps(nodes);
  for (const HostEstimate& e : pairwise) {
    if (e.confidence < confidence_threshold) continue;
    const size_t a = comps.root(e.source_idx), b = comps.root(e.target_idx);
    if (a != b) comps.join(a, b);
  }
  const size_t biggest = (size_t)(std::max_element(comps.count.begin(), comps.count.end()) - comps.count.begin());
  // membership is decided by the SOURCE node only, whatever the edge's own confidence (graph.cpp:94-99)
  std::vector<size_t> comp;
  for (size_t i = 0; i < pairwise.size(); ++i)
    if (comps.root(pairwise[i].source_idx) == biggest) {
      comp.push_back(i);
      if (in_component) (*in_component)[i] = 1;
    }

  // Kruskal, heaviest edge first; std::sort + std::greater keeps libstdc++'s tie order
  std::vector<WEdge> edges;
  for (size_t i : comp) edges.push_back(WEdge{pairwise[i].source_idx, pairwise[i].target_idx, pairwise[i].confidence});
  std::sort(edges.begin(), edges.end(), std::greater<WEdge>());
  size_t tree_nodes = 0;
  for (size_t i : comp) tree_nodes = std::max(tree_nodes, std::max(pairwise[i].source_idx + 1, pairwise[i].target_idx + 1));
  std::vector<std::vector<WEdge>> tree(tree_nodes);
  std::vector<size_t> degree(tree_nodes, 0);
  UnionFind forest(tree_nodes);
  for (const WEdge& e : edges) {
    const size_t a = forest.root(e.from), b = forest.root(e.to);
    if (a == b) continue;
    forest.join(a, b);
    tree[e.from].push_back(WEdge{e.from, e.to, e.weight});
    tree[e.to].push_back(WEdge{e.to, e.from, e.weight});
    degree[e.from]++;
    degree[e.to]++;
  }
  if (tree_edges)
    for (size_t v = 0; v < tree_nodes; ++v)
      for (const WEdge& e : tree[v]) tree_edges->push_back(std::make_pair((int)e.from, (int)e.to));

  // tree centre: nodes minimising the maximum hop distance to any leaf
  std::vector<size_t> ecc(tree_nodes, 0), hops;
  for (size_t leaf = 0; leaf < tree_nodes; ++leaf) {
    if (degree[leaf] != 1) continue;
    hops.assign(tree_nodes, 0);
    breadth_first(tree, leaf, [&](const WEdge& e) { hops[e.to] = hops[e.from] + 1; });
    for (size_t v = 0; v < tree_nodes; ++v) ecc[v] = std::max(ecc[v], hops[v]);
  }
  std::vector<size_t> centers;
  if (tree_nodes > 0) {
    const size_t best = *std::min_element(ecc.begin(), ecc.end());
    for (size_t v = 0; v < tree_nodes; ++v)
      if (ecc[v] == best) centers.push_back(v);
  }
  if (centers_out)
    for (size_t v : centers) centers_out->push_back((int)v);
  if (centers.empty()) return global;  // the reference indexes an empty vector here (undefined); all-zero is the defined stand-in

  const size_t ref = centers[0];
  if (reference_frame) *reference_frame = (int)ref;
  for (int i = 0; i < 16; ++i) global[ref][i] = (i % 5 == 0) ? 1.0f : 0.0f;
  breadth_first(tree, ref, [&](const WEdge& e) {
    // getTransform: first estimate matching (from,to) inverted, or (to,from) as is
    float t[16];
    for (int i = 0; i < 16; ++i) t[i] = 0.0f;
    for (size_t i : comp) {
      const HostEstimate& est = pairwise[i];
      if (est.source_idx == e.from && est.target_idx == e.to) { mat_inverse(est.T, t); break; }
      if (est.source_idx == e.to && est.target_idx == e.from) { for (int k = 0; k < 16; ++k) t[k] = est.T[k]; break; }
    }
    mat_mul(global[e.from].data(), t, global[e.to].data());
  });
  return global;
}

}  // namespace mm3d
