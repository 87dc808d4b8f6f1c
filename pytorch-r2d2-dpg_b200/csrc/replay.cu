// GPU-resident prioritized sequence replay shard (replaces LearnerReplayMemory, replay_memory.py:67-175).
//
// HBM layout (SoA, one "row" per stored env step incl. the n_step pad rows of actor.py:173):
//   obs_rows [cap,O]  act_rows [cap,A]  rew_rows [cap]  term_rows [cap]  state_rows [cap,4,2,H]
// Episodes occupy contiguous row ranges of a ring; FIFO eviction (replay_memory.py:148-152).
// Sum tree: one leaf per ROW (priority 0 for rows that are not valid sequence starts), fan-out 32:
// every node is the left-to-right fp32 sum of its 32 children (one 128-byte line), so the tree has
// ceil(log32(cap)) levels (5 for 2M rows) instead of 21 dependent loads of a binary tree, and the
// CUDA tree and its C restatement (oracle/sumtree_oracle.c) are bit-identical by construction.
#include <algorithm>
#include <deque>
#include <map>
#include <vector>

#include "replay.cuh"

namespace r2d2 {

namespace {

__global__ void __launch_bounds__(256) tree_sample_kernel(TreeView tv, const float* __restrict__ u, int batch,
                                                          long long* __restrict__ leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  const int top = tv.levels - 1;
  const float total = tv.lvl[top][0];
  float r = __fmul_rn(u[i], total);
  long long idx = 0;
  for (int l = top; l >= 1; --l) {
    const float4* ch4 = reinterpret_cast<const float4*>(tv.lvl[l - 1] + idx * TREE_K);
    float c[TREE_K];
#pragma unroll
    for (int k = 0; k < TREE_K / 4; ++k) {
      const float4 v = ch4[k];
      c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w;
    }
    int pick = -1;
#pragma unroll
    for (int k = 0; k < TREE_K; ++k) {
      if (pick < 0) {
        if (r < c[k]) pick = k;
        else r = __fsub_rn(r, c[k]);
      }
    }
    if (pick < 0) {  // rounding pushed the residual past the last child: take the last non-empty child
      float cl = 0.f;
#pragma unroll
      for (int k = 0; k < TREE_K; ++k)
        if (c[k] > 0.f) { pick = k; cl = c[k]; }
      if (pick < 0) pick = 0;
      r = __fmul_rn(cl, 0.99999994f);
    }
    idx = idx * TREE_K + pick;
  }
  leaf[i] = idx;
}

__device__ __forceinline__ float node_sum(const float* __restrict__ children) {
  const float4* ch4 = reinterpret_cast<const float4*>(children);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < TREE_K / 4; ++k) {
    const float4 v = ch4[k];
    s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
  }
  return s;
}

// single CTA: write the batch's leaves (highest batch index wins on duplicates, learner.py:136-139
// executes the writes in batch order), then refresh every ancestor level by level.  The leaf indices sit in shared
// memory: the last-writer test is a broadcast scan of the later entries (B^2 / 2 shared reads instead of global ones).
__global__ void __launch_bounds__(1024) tree_update_kernel(TreeView tv, const long long* __restrict__ leaf,
                                                           const float* __restrict__ prio, int batch) {
  extern __shared__ long long s_leaf[];
  for (int i = threadIdx.x; i < batch; i += blockDim.x) s_leaf[i] = leaf[i];
  __syncthreads();
  for (int i = threadIdx.x; i < batch; i += blockDim.x) {
    const long long li = s_leaf[i];
    bool winner = true;
    for (int j = i + 1; j < batch; ++j)
      if (s_leaf[j] == li) { winner = false; break; }
    if (winner) tv.lvl[0][li] = prio[i];
  }
  __syncthreads();
  long long div = TREE_K;
  for (int l = 1; l < tv.levels; ++l) {
    for (int i = threadIdx.x; i < batch; i += blockDim.x) {
      const long long node = s_leaf[i] / div;
      tv.lvl[l][node] = node_sum(tv.lvl[l - 1] + node * TREE_K);
    }
    __syncthreads();
    div *= TREE_K;
  }
}

__global__ void __launch_bounds__(256) tree_recompute_range_kernel(TreeView tv, int level, long long first,
                                                                   long long count) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= count) return;
  const long long node = first + i;
  tv.lvl[level][node] = node_sum(tv.lvl[level - 1] + node * TREE_K);
}

// One launch gathers the whole time-major batch (replay_memory.py:116-133: slice, transpose, stack, H2D in the
// reference).  A warp owns one (t, b) pair: source row leaf[b] + t; it copies the obs row and the act row with
// lane-contiguous accesses (16-byte vectors when the row width is a multiple of 4 floats: a row start is then 16-byte
// aligned in both the shard and the batch), lane 0 moves the reward / terminal scalars.  Tasks >= T*B move the stored
// recurrent states: task T*B + nh*B + b copies state_rows[leaf[b]][nh][:] to out[nh][b][:].  No division per element.
struct GatherParams {
  const float *obs_rows, *act_rows, *rew_rows, *term_rows, *state_rows;
  const long long* leaf;
  float *obs, *act, *rew, *term, *states;
  int T, B, O, A, H;
};

__device__ __forceinline__ void warp_copy_row(const float* __restrict__ src, float* __restrict__ dst, int n, int lane) {
  if ((n & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int k = lane; k < (n >> 2); k += 32) d4[k] = __ldg(s4 + k);
  } else {
    for (int k = lane; k < n; k += 32) dst[k] = __ldg(src + k);
  }
}

__global__ void __launch_bounds__(256) gather_batch_kernel(GatherParams g) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long row_tasks = (long long)g.T * g.B;
  const long long total = row_tasks + (g.states ? (long long)8 * g.B : 0);
  int t = (int)(warp0 / g.B), b = (int)(warp0 % g.B);                    // one division per warp, then carried
  const int dt = (int)(n_warps / g.B), db = (int)(n_warps % g.B);
  for (long long task = warp0; task < total; task += n_warps) {
    if (task < row_tasks) {
      const long long r = g.leaf[b] + t;
      const long long o = (long long)t * g.B + b;
      if (g.obs) warp_copy_row(g.obs_rows + r * g.O, g.obs + o * g.O, g.O, lane);
      if (g.act) warp_copy_row(g.act_rows + r * g.A, g.act + o * g.A, g.A, lane);
      if (lane == 0) {
        if (g.rew) g.rew[o] = __ldg(g.rew_rows + r);
        if (g.term) g.term[o] = __ldg(g.term_rows + r);
      }
    } else {
      warp_copy_row(g.state_rows + (g.leaf[b] * 8 + (t - g.T)) * g.H, g.states + ((long long)(t - g.T) * g.B + b) * g.H, g.H, lane);
    }
    t += dt; b += db;
    if (b >= g.B) { b -= g.B; ++t; }
  }
}

int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

struct Episode {
  long long row_start;
  int n_rows;
  int n_starts;
  long long serial;
};

struct Replay {
  r2d2_replay_config cfg;
  int rows_per_window;
  float *obs_rows = nullptr, *act_rows = nullptr, *rew_rows = nullptr, *term_rows = nullptr, *state_rows = nullptr;
  std::vector<float*> level_alloc;
  TreeView tv;
  std::deque<Episode> episodes;
  std::map<long long, long long> by_row;  // row_start -> serial
  long long next_serial = 0;
  long long head = 0;
  long long sequence_counter = 0;
  long long rows_used = 0;
  long long evicted_total = 0;
};

static int recompute_ancestors(Replay* r, long long first_leaf, long long n_leaves, cudaStream_t stream) {
  long long lo = first_leaf, hi = first_leaf + n_leaves - 1;
  for (int l = 1; l < r->tv.levels; ++l) {
    lo /= TREE_K;
    hi /= TREE_K;
    const long long count = hi - lo + 1;
    tree_recompute_range_kernel<<<(int)((count + 255) / 256), 256, 0, stream>>>(r->tv, l, lo, count);
    count_launch();
  }
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int replay_create(Replay** out, const r2d2_replay_config* cfg) {
  R2D2_REQUIRE(out && cfg, "null");
  R2D2_REQUIRE(cfg->obs_size > 0 && cfg->n_actions > 0 && cfg->hidden > 0, "sizes");
  R2D2_REQUIRE(cfg->capacity_rows > 0, "capacity_rows");
  Replay* r = new Replay();
  r->cfg = *cfg;
  r->rows_per_window = cfg->burn_in + cfg->learning + cfg->n_step;
  const long long cap = cfg->capacity_rows;
  auto dmalloc = [&](float** p, long long n) -> int {
    R2D2_CUDA_TRY(cudaMalloc(p, sizeof(float) * (size_t)n));
    R2D2_CUDA_TRY(cudaMemset(*p, 0, sizeof(float) * (size_t)n));
    return R2D2_OK;
  };
  int rc = R2D2_OK;
  if ((rc = dmalloc(&r->obs_rows, cap * cfg->obs_size)) || (rc = dmalloc(&r->act_rows, cap * cfg->n_actions)) ||
      (rc = dmalloc(&r->rew_rows, cap)) || (rc = dmalloc(&r->term_rows, cap)) ||
      (rc = dmalloc(&r->state_rows, cap * 8 * cfg->hidden))) {
    delete r;
    return rc;
  }
  // levels: n[0] = leaves, n[l+1] = ceil(n[l]/32), last level has one node (the total)
  long long n = (cap + TREE_K - 1) / TREE_K * TREE_K;
  int levels = 0;
  while (true) {
    R2D2_REQUIRE(levels < TREE_MAX_LEVELS, "tree too deep");
    const long long parents = (n + TREE_K - 1) / TREE_K;
    const long long alloc = (n > 1) ? parents * TREE_K : TREE_K;
    float* p = nullptr;
    if ((rc = dmalloc(&p, alloc))) { delete r; return rc; }
    r->level_alloc.push_back(p);
    r->tv.lvl[levels] = p;
    r->tv.n[levels] = n;
    ++levels;
    if (n == 1) break;
    n = parents;
  }
  r->tv.levels = levels;
  *out = r;
  return R2D2_OK;
}

int replay_destroy(Replay* r) {
  if (!r) return R2D2_OK;
  cudaFree(r->obs_rows); cudaFree(r->act_rows); cudaFree(r->rew_rows); cudaFree(r->term_rows); cudaFree(r->state_rows);
  for (float* p : r->level_alloc) cudaFree(p);
  delete r;
  return R2D2_OK;
}

// Ranges of leaves whose ancestors must be refreshed; merged and recomputed once per ingest call.
typedef std::vector<std::pair<long long, long long>> RangeList;   // (first leaf, count)

static int refresh_ranges(Replay* r, RangeList& ranges, cudaStream_t stream) {
  if (ranges.empty()) return R2D2_OK;
  std::sort(ranges.begin(), ranges.end());
  long long lo = ranges[0].first, hi = ranges[0].first + ranges[0].second;
  for (size_t i = 1; i <= ranges.size(); ++i) {
    // merge ranges that share a parent node (distance < fan-out): one launch chain instead of two
    if (i < ranges.size() && ranges[i].first <= hi + TREE_K) {
      hi = std::max(hi, ranges[i].first + ranges[i].second);
      continue;
    }
    R2D2_TRY(recompute_ancestors(r, lo, hi - lo, stream));
    if (i < ranges.size()) { lo = ranges[i].first; hi = ranges[i].first + ranges[i].second; }
  }
  ranges.clear();
  return R2D2_OK;
}

static int evict_front(Replay* r, cudaStream_t stream, RangeList* deferred) {
  const Episode e = r->episodes.front();
  r->episodes.pop_front();
  r->by_row.erase(e.row_start);
  // replay_memory.py:149: the counter drops by len(episode) - sequence_length (sic: not the amount added at :147)
  r->sequence_counter -= e.n_rows - (r->cfg.burn_in + r->cfg.learning);
  r->rows_used -= e.n_rows;
  ++r->evicted_total;
  if (e.n_starts > 0) {
    R2D2_CUDA_TRY(cudaMemsetAsync(r->tv.lvl[0] + e.row_start, 0, sizeof(float) * e.n_starts, stream));
    if (deferred) deferred->push_back({e.row_start, (long long)e.n_starts});
    else R2D2_TRY(recompute_ancestors(r, e.row_start, e.n_starts, stream));
  }
  return R2D2_OK;
}

// Ring placement of an episode of n_rows rows: wrap when the tail gap is too small, then evict - oldest first, the
// reference's FIFO order (replay_memory.py:148-152) - until no live episode overlaps [start, start + n_rows).  After a
// wrap the oldest episode may sit at the TAIL of the ring while a younger one occupies the head that is about to be
// overwritten: evicting from the front until the overlap is gone removes both (the earlier loop stopped at
// the first non-overlapping front episode and then failed the overlap check).
static int place_episode(Replay* r, int n_rows, cudaStream_t stream, RangeList* deferred, long long* start_out) {
  R2D2_REQUIRE(n_rows <= r->cfg.capacity_rows, "episode larger than the ring");
  if (r->head + n_rows > r->cfg.capacity_rows) r->head = 0;  // wrap: the tail gap stays unused
  const long long start = r->head, end = start + n_rows;
  auto overlaps = [&]() {
    for (const Episode& e : r->episodes)
      if (e.row_start < end && start < e.row_start + e.n_rows) return true;
    return false;
  };
  while (!r->episodes.empty() && overlaps()) R2D2_TRY(evict_front(r, stream, deferred));
  *start_out = start;
  return R2D2_OK;
}

static void commit_episode(Replay* r, long long start, int n_rows, int n_starts) {
  Episode e{start, n_rows, n_starts, r->next_serial++};
  r->episodes.push_back(e);
  r->by_row[start] = e.serial;
  r->head = start + n_rows;
  r->rows_used += n_rows;
  r->sequence_counter += n_rows - (r->rows_per_window - 1);  // replay_memory.py:147
}

int replay_add_episode(Replay* r, const float* obs, const float* act, const float* rew, const float* term,
                       const float* states, int n_rows, int n_state_rows, const float* priority, int n_starts,
                       cudaStream_t stream) {
  R2D2_REQUIRE(r && obs && act && rew && term && states, "null");
  R2D2_REQUIRE(n_rows >= r->rows_per_window, "episode shorter than one window");
  R2D2_REQUIRE(n_starts >= 0 && n_starts <= n_rows - r->rows_per_window + 1, "n_starts exceeds valid window starts");
  R2D2_REQUIRE(n_state_rows >= n_starts && n_state_rows <= n_rows, "state rows");
  R2D2_REQUIRE(n_starts == 0 || priority, "priority");
  const int O = r->cfg.obs_size, A = r->cfg.n_actions, H = r->cfg.hidden;
  RangeList ranges;
  long long start = 0;
  R2D2_TRY(place_episode(r, n_rows, stream, &ranges, &start));
  R2D2_CUDA_TRY(cudaMemcpyAsync(r->obs_rows + start * O, obs, sizeof(float) * (size_t)n_rows * O, cudaMemcpyHostToDevice, stream));
  R2D2_CUDA_TRY(cudaMemcpyAsync(r->act_rows + start * A, act, sizeof(float) * (size_t)n_rows * A, cudaMemcpyHostToDevice, stream));
  R2D2_CUDA_TRY(cudaMemcpyAsync(r->rew_rows + start, rew, sizeof(float) * (size_t)n_rows, cudaMemcpyHostToDevice, stream));
  R2D2_CUDA_TRY(cudaMemcpyAsync(r->term_rows + start, term, sizeof(float) * (size_t)n_rows, cudaMemcpyHostToDevice, stream));
  R2D2_CUDA_TRY(cudaMemcpyAsync(r->state_rows + start * 8 * H, states, sizeof(float) * (size_t)n_state_rows * 8 * H,
                                cudaMemcpyHostToDevice, stream));
  if (n_state_rows < n_rows)
    R2D2_CUDA_TRY(cudaMemsetAsync(r->state_rows + (start + n_state_rows) * 8 * H, 0,
                                  sizeof(float) * (size_t)(n_rows - n_state_rows) * 8 * H, stream));
  if (n_starts > 0)
    R2D2_CUDA_TRY(cudaMemcpyAsync(r->tv.lvl[0] + start, priority, sizeof(float) * (size_t)n_starts,
                                  cudaMemcpyHostToDevice, stream));
  if (n_rows > n_starts)
    R2D2_CUDA_TRY(cudaMemsetAsync(r->tv.lvl[0] + start + n_starts, 0, sizeof(float) * (size_t)(n_rows - n_starts), stream));
  ranges.push_back({start, (long long)n_rows});
  commit_episode(r, start, n_rows, n_starts);
  while (r->cfg.max_sequences > 0 && r->sequence_counter > r->cfg.max_sequences && r->episodes.size() > 1)
    R2D2_TRY(evict_front(r, stream, &ranges));
  R2D2_TRY(refresh_ranges(r, ranges, stream));
  R2D2_CUDA_TRY(cudaStreamSynchronize(stream));  // host buffers may be released by the caller
  return R2D2_OK;
}

// One actor file at a time (LearnerReplayMemory.load, replay_memory.py:138-157): every episode of the file is appended,
// THEN the oldest episodes are dropped while the sequence counter exceeds the cap - the reference's order.  The rows
// of all episodes arrive packed ([R, *] with R = sum of n_rows; recurrent states zero-padded to R rows, leaf
// priorities already expanded to one value per row): contiguous runs in the ring are one copy per tensor, the sum tree
// is refreshed once over the merged touched ranges, and there is one stream synchronisation per file.
int replay_add_episodes(Replay* r, int n_episodes, const int* n_rows, const int* n_starts, const float* obs,
                        const float* act, const float* rew, const float* term, const float* states,
                        const float* leaf_prio, long long* row_start_out, long long* n_evicted_out,
                        long long* sequence_counter_out, cudaStream_t stream) {
  R2D2_REQUIRE(r && n_episodes >= 0 && (n_episodes == 0 || (n_rows && n_starts && obs && act && rew && term && states && leaf_prio)),
               "null");
  const int O = r->cfg.obs_size, A = r->cfg.n_actions, H = r->cfg.hidden;
  for (int e = 0; e < n_episodes; ++e) {
    R2D2_REQUIRE(n_rows[e] >= r->rows_per_window, "episode shorter than one window");
    R2D2_REQUIRE(n_starts[e] >= 0 && n_starts[e] <= n_rows[e] - r->rows_per_window + 1, "n_starts exceeds valid window starts");
    R2D2_REQUIRE(n_rows[e] <= r->cfg.capacity_rows, "episode larger than the ring");
  }
  const long long evicted0 = r->evicted_total;
  RangeList ranges;
  long long src = 0;                    // first packed row of the current run
  long long run_start = -1, run_rows = 0;
  auto flush = [&]() -> int {
    if (run_rows == 0) return R2D2_OK;
    const size_t n = (size_t)run_rows;
    R2D2_CUDA_TRY(cudaMemcpyAsync(r->obs_rows + run_start * O, obs + src * O, sizeof(float) * n * O, cudaMemcpyHostToDevice, stream));
    R2D2_CUDA_TRY(cudaMemcpyAsync(r->act_rows + run_start * A, act + src * A, sizeof(float) * n * A, cudaMemcpyHostToDevice, stream));
    R2D2_CUDA_TRY(cudaMemcpyAsync(r->rew_rows + run_start, rew + src, sizeof(float) * n, cudaMemcpyHostToDevice, stream));
    R2D2_CUDA_TRY(cudaMemcpyAsync(r->term_rows + run_start, term + src, sizeof(float) * n, cudaMemcpyHostToDevice, stream));
    R2D2_CUDA_TRY(cudaMemcpyAsync(r->state_rows + run_start * 8 * H, states + src * 8 * H, sizeof(float) * n * 8 * H,
                                  cudaMemcpyHostToDevice, stream));
    R2D2_CUDA_TRY(cudaMemcpyAsync(r->tv.lvl[0] + run_start, leaf_prio + src, sizeof(float) * n, cudaMemcpyHostToDevice, stream));
    ranges.push_back({run_start, run_rows});
    src += run_rows;
    run_rows = 0;
    return R2D2_OK;
  };
  for (int e = 0; e < n_episodes; ++e) {
    long long start = 0;
    R2D2_TRY(place_episode(r, n_rows[e], stream, &ranges, &start));
    if (run_rows > 0 && start != run_start + run_rows) R2D2_TRY(flush());   // the ring wrapped: new run
    if (run_rows == 0) run_start = start;
    run_rows += n_rows[e];
    commit_episode(r, start, n_rows[e], n_starts[e]);
    if (row_start_out) row_start_out[e] = start;
  }
  R2D2_TRY(flush());
  while (r->cfg.max_sequences > 0 && r->sequence_counter > r->cfg.max_sequences && !r->episodes.empty())
    R2D2_TRY(evict_front(r, stream, &ranges));                               // replay_memory.py:148-152
  R2D2_TRY(refresh_ranges(r, ranges, stream));
  R2D2_CUDA_TRY(cudaStreamSynchronize(stream));  // host buffers may be released by the caller
  if (n_evicted_out) *n_evicted_out = r->evicted_total - evicted0;
  if (sequence_counter_out) *sequence_counter_out = r->sequence_counter;
  return R2D2_OK;
}

int replay_gather(Replay* r, const long long* leaf_idx, int batch, float* obs, float* act, float* rew, float* term,
                  float* states, cudaStream_t stream) {
  R2D2_REQUIRE(r && leaf_idx && batch > 0, "args");
  const int T = r->rows_per_window, O = r->cfg.obs_size, A = r->cfg.n_actions, H = r->cfg.hidden;
  if (obs || act || rew || term || states) {
    GatherParams g;
    g.obs_rows = r->obs_rows; g.act_rows = r->act_rows; g.rew_rows = r->rew_rows; g.term_rows = r->term_rows;
    g.state_rows = r->state_rows; g.leaf = leaf_idx;
    g.obs = obs; g.act = act; g.rew = rew; g.term = term; g.states = states;
    g.T = T; g.B = batch; g.O = O; g.A = A; g.H = H;
    const long long tasks = (long long)T * batch + (states ? (long long)8 * batch : 0);
    gather_batch_kernel<<<grid_for(tasks * 32), 256, 0, stream>>>(g);
    count_launch();
  }
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int replay_sample(Replay* r, const float* u, int batch, long long* leaf_idx, float* obs, float* act, float* rew,
                  float* term, float* states, cudaStream_t stream) {
  R2D2_REQUIRE(r && u && leaf_idx && batch > 0, "args");
  R2D2_REQUIRE(!r->episodes.empty(), "replay is empty");
  tree_sample_kernel<<<ceil_div(batch, 256), 256, 0, stream>>>(r->tv, u, batch, leaf_idx);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return replay_gather(r, leaf_idx, batch, obs, act, rew, term, states, stream);
}

int replay_update_priorities(Replay* r, const long long* leaf_idx, const float* prio, int batch, cudaStream_t stream) {
  R2D2_REQUIRE(r && leaf_idx && prio && batch > 0, "args");
  R2D2_REQUIRE(batch <= 5120, "priority batch larger than the shared-memory index table (40 KB)");
  tree_update_kernel<<<1, 1024, sizeof(long long) * (size_t)batch, stream>>>(r->tv, leaf_idx, prio, batch);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int replay_stats(Replay* r, r2d2_replay_stats_t* out, cudaStream_t stream) {
  R2D2_REQUIRE(r && out, "args");
  float total = 0.f;
  R2D2_CUDA_TRY(cudaMemcpyAsync(&total, r->tv.lvl[r->tv.levels - 1], sizeof(float), cudaMemcpyDeviceToHost, stream));
  R2D2_CUDA_TRY(cudaStreamSynchronize(stream));
  out->n_episodes = (long long)r->episodes.size();
  out->n_rows_used = r->rows_used;
  out->sequence_counter = r->sequence_counter;
  out->capacity_rows = r->cfg.capacity_rows;
  out->tree_levels = r->tv.levels;
  long long nodes = 0;
  for (int l = 0; l < r->tv.levels; ++l) nodes += r->tv.n[l];
  out->tree_nodes = nodes;
  out->last_row_start = r->episodes.empty() ? -1 : r->episodes.back().row_start;
  out->total_priority = total;
  return R2D2_OK;
}

int replay_decode(Replay* r, const long long* leaf_host, int n, long long* episode_index, long long* sequence_index) {
  R2D2_REQUIRE(r && leaf_host && episode_index && sequence_index, "args");
  const long long front_serial = r->episodes.empty() ? 0 : r->episodes.front().serial;
  for (int i = 0; i < n; ++i) {
    episode_index[i] = sequence_index[i] = -1;
    auto it = r->by_row.upper_bound(leaf_host[i]);
    if (it == r->by_row.begin()) continue;
    --it;
    const Episode& e = r->episodes[(size_t)(it->second - front_serial)];
    if (leaf_host[i] < e.row_start + e.n_rows) {
      episode_index[i] = it->second - front_serial;
      sequence_index[i] = leaf_host[i] - e.row_start;
    }
  }
  return R2D2_OK;
}

int replay_tree_level(Replay* r, int level, const float** dev_ptr, long long* n) {
  R2D2_REQUIRE(r && level >= 0 && level < r->tv.levels, "level");
  *dev_ptr = r->tv.lvl[level];
  *n = r->tv.n[level];
  return R2D2_OK;
}

}  // namespace r2d2
