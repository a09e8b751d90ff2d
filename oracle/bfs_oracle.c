/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the reference's unweighted shortest paths:
 * a queue-based BFS from one node (graphembed/graphembed/pyx/impl/precision.cpp:44-63, the in-tree statement of
 * what networkit's APSP computes for compute_graph_pdists, graphembed/graphembed/data/graph.py:66-87).
 * Parity: pinned against scipy.sparse.csgraph.shortest_path / networkx in tests/test_oracle_bfs.py and against
 * the golden hop counts in tests/golden/training_run_f64.npz. */
#include <stdint.h>
#include <stdlib.h>

/* dist[v] = hop count from `src`, -1 if unreachable.  CSR lists, for every node, the nodes it is adjacent to. */
void bfs_from(const int32_t* rowptr, const int32_t* colidx, int32_t n, int32_t src, int32_t* dist, int32_t* queue) {
  for (int32_t i = 0; i < n; ++i) dist[i] = -1;
  int32_t head = 0, tail = 0;
  queue[tail++] = src;
  dist[src] = 0;
  while (head < tail) {
    int32_t node = queue[head++];
    for (int32_t e = rowptr[node]; e < rowptr[node + 1]; ++e) {
      int32_t nb = colidx[e];
      if (dist[nb] == -1) {
        dist[nb] = dist[node] + 1;
        queue[tail++] = nb;
      }
    }
  }
}

/* out[s*n + v] for every source in `sources` */
int bfs_many(const int32_t* rowptr, const int32_t* colidx, int32_t n, const int32_t* sources, int32_t n_sources,
             int32_t* out) {
  int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  if (!queue) return -1;
  for (int32_t s = 0; s < n_sources; ++s) bfs_from(rowptr, colidx, n, sources[s], out + (size_t)s * n, queue);
  free(queue);
  return 0;
}
