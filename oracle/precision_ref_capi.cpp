// TEST INFRASTRUCTURE ONLY -- extern "C" doorway to the REFERENCE's FastPrecision (graphembed/pyx/impl/precision.hpp,
// compiled from /root/reference by oracle/Makefile into oracle/_ref/libprecision_ref.so): builds the AdjList the Cython
// wrapper builds (pyx/precision.pyx:60-72) from a CSR adjacency and forwards to the reference's methods.
#include <cstring>
#include "graphembed/pyx/impl/precision.hpp"

using graphembed::AdjList;
using graphembed::FastPrecision;

extern "C" {
void* fpref_create(int n, const int* rowptr, const int* colidx) {
  AdjList adj(n);
  for (int u = 0; u < n; ++u)
    for (int e = rowptr[u]; e < rowptr[u + 1]; ++e) adj[u].insert(colidx[e]);
  return new FastPrecision(std::move(adj));
}
void fpref_destroy(void* h) { delete static_cast<FastPrecision*>(h); }
int fpref_nodes_per_layer(void* h, int* out, int cap) {
  auto v = static_cast<FastPrecision*>(h)->NodesPerLayer();
  int k = (int)v.size() < cap ? (int)v.size() : cap;
  std::memcpy(out, v.data(), k * sizeof(int));
  return (int)v.size();
}
double fpref_map_f64(void* h, const double* mp) { return static_cast<FastPrecision*>(h)->MeanAveragePrecision(mp); }
double fpref_map_f32(void* h, const float* mp) { return static_cast<FastPrecision*>(h)->MeanAveragePrecision(mp); }
static int put(const graphembed::StatsR& r, double* means, double* stds, int cap) {
  int k = (int)r.means.size() < cap ? (int)r.means.size() : cap;
  std::memcpy(means, r.means.data(), k * sizeof(double));
  std::memcpy(stds, r.stds.data(), k * sizeof(double));
  return (int)r.means.size();
}
int fpref_layer_f1_f64(void* h, const double* mp, long min_deg, long max_deg, double* means, double* stds, int cap) {
  return put(static_cast<FastPrecision*>(h)->LayerMeanF1Scores(mp, 1, (size_t)min_deg, (size_t)max_deg), means, stds, cap);
}
int fpref_layer_f1_f32(void* h, const float* mp, long min_deg, long max_deg, double* means, double* stds, int cap) {
  return put(static_cast<FastPrecision*>(h)->LayerMeanF1Scores(mp, 1, (size_t)min_deg, (size_t)max_deg), means, stds, cap);
}
int fpref_layer_avg_f1_f64(void* h, const double* mp, double* means, double* stds, int cap) {
  return put(static_cast<FastPrecision*>(h)->LayerMeanAverageF1Scores(mp, 1), means, stds, cap);
}
}
