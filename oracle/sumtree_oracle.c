/* CPU restatement of the 32-ary fp32 sum tree of pytorch-r2d2-dpg_b200/csrc/replay.cu.
 *
 * TEST INFRASTRUCTURE: linked only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg (through oracle/sumtree.py).  The product path never calls it.
 *
 * What it pins: the reference has NO tree - LearnerReplayMemory draws an episode proportionally to
 * total_priority and then a sequence proportionally to priority[ep] (replay_memory.py:95-114), i.e.
 * P(ep, seq) = priority[ep][seq] / sum(all priorities).  A proportional draw over the flat array of
 * sequence-start priorities is the same distribution; this file defines, in plain C, exactly which
 * index that draw returns for a given uniform u so the CUDA kernel can be checked bit for bit
 * (same tree shape, same left-to-right fp32 summation order, same descent rule).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile); x86-64 SSE float arithmetic.
 */
#include <stdlib.h>
#include <string.h>

#define K 32
#define MAX_LEVELS 8

typedef struct {
  int levels;
  long long n[MAX_LEVELS];
  float* lvl[MAX_LEVELS];
} tree_t;

static float node_sum(const float* c) {
  float s = 0.0f;
  for (int k = 0; k < K; ++k) s = s + c[k]; /* left to right, one fp32 add each */
  return s;
}

void* st_create(long long capacity) {
  tree_t* t = (tree_t*)calloc(1, sizeof(tree_t));
  long long n = (capacity + K - 1) / K * K;
  int levels = 0;
  for (;;) {
    long long parents = (n + K - 1) / K;
    long long alloc = (n > 1) ? parents * K : K;
    t->lvl[levels] = (float*)calloc((size_t)alloc, sizeof(float));
    t->n[levels] = n;
    ++levels;
    if (n == 1) break;
    n = parents;
  }
  t->levels = levels;
  return t;
}

void st_destroy(void* h) {
  tree_t* t = (tree_t*)h;
  for (int l = 0; l < t->levels; ++l) free(t->lvl[l]);
  free(t);
}

int st_levels(void* h) { return ((tree_t*)h)->levels; }
long long st_level_size(void* h, int l) { return ((tree_t*)h)->n[l]; }
const float* st_level_ptr(void* h, int l) { return ((tree_t*)h)->lvl[l]; }
float st_total(void* h) { tree_t* t = (tree_t*)h; return t->lvl[t->levels - 1][0]; }

/* leaves[first .. first+count) = src (or 0 when src == NULL), then refresh the ancestors of the range */
void st_set_range(void* h, long long first, long long count, const float* src) {
  tree_t* t = (tree_t*)h;
  if (count <= 0) return;
  if (src) memcpy(t->lvl[0] + first, src, (size_t)count * sizeof(float));
  else memset(t->lvl[0] + first, 0, (size_t)count * sizeof(float));
  long long lo = first, hi = first + count - 1;
  for (int l = 1; l < t->levels; ++l) {
    lo /= K; hi /= K;
    for (long long node = lo; node <= hi; ++node) t->lvl[l][node] = node_sum(t->lvl[l - 1] + node * K);
  }
}

/* batch write-back: on duplicate leaves the highest batch index wins (the python loop at
 * learner.py:136-139 writes in batch order), then every ancestor is recomputed from its children */
void st_update_batch(void* h, const long long* leaf, const float* prio, int n) {
  tree_t* t = (tree_t*)h;
  for (int i = 0; i < n; ++i) t->lvl[0][leaf[i]] = prio[i];
  long long div = K;
  for (int l = 1; l < t->levels; ++l) {
    for (int i = 0; i < n; ++i) {
      long long node = leaf[i] / div;
      t->lvl[l][node] = node_sum(t->lvl[l - 1] + node * K);
    }
    div *= K;
  }
}

/* proportional draw: residual r = u * total walks down; at a node the children are scanned left to
 * right, the first child with r < c is taken, otherwise r -= c.  If rounding pushes r past the last
 * child, the last non-empty child is taken with r just below its sum. */
void st_sample(void* h, const float* u, int n, long long* out) {
  tree_t* t = (tree_t*)h;
  const int top = t->levels - 1;
  const float total = t->lvl[top][0];
  for (int i = 0; i < n; ++i) {
    float r = u[i] * total;
    long long idx = 0;
    for (int l = top; l >= 1; --l) {
      const float* c = t->lvl[l - 1] + idx * K;
      int pick = -1;
      for (int k = 0; k < K; ++k) {
        if (r < c[k]) { pick = k; break; }
        r = r - c[k];
      }
      if (pick < 0) {
        float cl = 0.0f;
        for (int k = 0; k < K; ++k) if (c[k] > 0.0f) { pick = k; cl = c[k]; }
        if (pick < 0) pick = 0;
        r = cl * 0.99999994f;
      }
      idx = idx * K + pick;
    }
    out[i] = idx;
  }
}
