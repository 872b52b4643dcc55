// Host-side planner of the tensor-core SCAN path: packs variable-length captions into
// 128-row word tiles (the UMMA M dimension).
//
// Why: the reference slices every caption to its true length before attention
// (Objectives.py:340-341), so the tensor-core formulation must never spend MMA rows on
// padding, and the epilogue's per-caption reductions (l2norm over the caption's words,
// LSE/mean over words) must stay inside one warp.  Captions of <= 32 words are therefore
// bin-packed (best-fit decreasing) into 32-row quarters -- one quarter = the 32 TMEM lanes
// one epilogue warp can read -- and four quarters make a tile.  A caption longer than 32
// words gets a tile of its own, flagged `long`, whose reductions cross warps through
// shared memory.  The scoring order of captions is free (every (image, caption) pair is
// independent), so packing only permutes work; results land at the original caption index.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/itr_b200.h"

namespace itr {
int fail(int code, const char* fmt, ...);
}

namespace {

struct Quarter {
  std::vector<int> caps;   // caption ids in placement order
  int used = 0;
};

struct Plan {
  std::vector<int> long_caps;        // one tile each
  std::vector<Quarter> quarters;     // four per tile, in creation order
  int n_tiles() const { return (int)long_caps.size() + ((int)quarters.size() + 3) / 4; }
};

int build_plan(const int32_t* lens, int n_cap, Plan& plan) {
  std::vector<int> order;
  order.reserve(n_cap);
  for (int c = 0; c < n_cap; ++c) {
    if (lens[c] < 1 || lens[c] > ITR_TILE_WORDS)
      return itr::fail(ITR_ERR_INVALID, "caption %d has length %d; the tensor-core path supports 1..%d words", c,
                       (int)lens[c], ITR_TILE_WORDS);
    if (lens[c] > 32) plan.long_caps.push_back(c);
    else order.push_back(c);
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lens[a] > lens[b]; });
  // open[r] = quarters with exactly r free rows (LIFO keeps placement deterministic)
  std::vector<std::vector<int>> open(33);
  for (int c : order) {
    const int len = lens[c];
    int q = -1;
    for (int r = len; r <= 32 && q < 0; ++r) {
      if (!open[r].empty()) { q = open[r].back(); open[r].pop_back(); }
    }
    if (q < 0) { q = (int)plan.quarters.size(); plan.quarters.emplace_back(); }
    Quarter& Q = plan.quarters[q];
    Q.caps.push_back(c);
    Q.used += len;
    if (Q.used < 32) open[32 - Q.used].push_back(q);
  }
  return ITR_OK;
}

}  // namespace

extern "C" int itr_scan_plan_max_tiles(const int32_t* cap_lens_host, int n_cap) {
  if (!cap_lens_host || n_cap < 0) { itr::fail(ITR_ERR_INVALID, "itr_scan_plan_max_tiles: bad arguments"); return -ITR_ERR_INVALID; }
  Plan plan;
  int rc = build_plan(cap_lens_host, n_cap, plan);
  if (rc) return -rc;
  return plan.n_tiles();
}

extern "C" int itr_scan_plan_words(const int32_t* cap_lens_host, int n_cap, int32_t* row_meta_host, int* n_tiles) {
  if (!cap_lens_host || !row_meta_host || !n_tiles || n_cap < 0)
    return itr::fail(ITR_ERR_INVALID, "itr_scan_plan_words: bad arguments");
  Plan plan;
  int rc = build_plan(cap_lens_host, n_cap, plan);
  if (rc) return rc;
  const int T = plan.n_tiles();
  *n_tiles = T;
  // default: every row is padding and its own one-lane segment
  for (int t = 0; t < T; ++t)
    for (int row = 0; row < ITR_TILE_WORDS; ++row) {
      int32_t* m = row_meta_host + 4 * ((size_t)t * ITR_TILE_WORDS + row);
      const int lane = row & 31;
      m[0] = -1; m[1] = 0; m[2] = lane | (lane << 8); m[3] = 0;
    }
  int t = 0;
  for (int c : plan.long_caps) {           // long tiles first
    const int len = cap_lens_host[c];
    for (int row = 0; row < ITR_TILE_WORDS; ++row) {
      int32_t* m = row_meta_host + 4 * ((size_t)t * ITR_TILE_WORDS + row);
      m[2] = (0) | (31 << 8) | (1 << 16);  // reductions span the whole tile
      if (row < len) { m[0] = c; m[1] = row; m[3] = len; }
    }
    ++t;
  }
  for (size_t q = 0; q < plan.quarters.size(); ++q) {
    const size_t tile = (size_t)t + q / 4;
    int lane = 0;
    for (int c : plan.quarters[q].caps) {
      const int len = cap_lens_host[c];
      for (int j = 0; j < len; ++j) {
        int32_t* m = row_meta_host + 4 * (tile * ITR_TILE_WORDS + (q % 4) * 32 + lane + j);
        m[0] = c; m[1] = j; m[2] = lane | ((lane + len - 1) << 8); m[3] = len;
      }
      lane += len;
    }
  }
  return ITR_OK;
}

// Items of the ground-truth pre-pass of the fused evaluation (itr_scan_t2i_gt_thresholds_bf16).  A word tile "needs"
// an image tile when one of its packed captions belongs to one of the tile's images (global caption cap_offset + c
// belongs to image (cap_offset + c) / caps_per_img; images come in tiles of ITR_TILE_IMAGES).  The CTA pair of the kernel
// scores TWO word tiles against one image tile per item, so the word tiles that need image tile n are paired up:
// items_host receives int32 quadruples (word tile of the leader, word tile of the peer or n_tiles for none, image
// tile, 0), sorted by image tile, or may be NULL to just count them; returns the count through *n_items.
extern "C" int itr_scan_plan_gt_items(const int32_t* row_meta_host, int n_tiles, int cap_offset, int caps_per_img, int n_img,
                                      int32_t* items_host, int max_items, int* n_items) {
  if (!row_meta_host || !n_items || n_tiles < 0 || cap_offset < 0 || caps_per_img < 1 || n_img < 0)
    return itr::fail(ITR_ERR_INVALID, "itr_scan_plan_gt_items: bad arguments");
  const int n_it = (n_img + ITR_TILE_IMAGES - 1) / ITR_TILE_IMAGES;
  std::vector<std::pair<int, int>> needs;      // (image tile, word tile)
  std::vector<int> seen;
  for (int t = 0; t < n_tiles; ++t) {
    seen.clear();
    for (int row = t * ITR_TILE_WORDS; row < (t + 1) * ITR_TILE_WORDS; ++row) {
      const int32_t* m = row_meta_host + 4 * (size_t)row;
      if (m[0] < 0 || m[1] != 0) continue;      // one look per caption: its first word
      const long long img = ((long long)cap_offset + m[0]) / caps_per_img;
      if (img >= n_img) continue;
      seen.push_back((int)(img / ITR_TILE_IMAGES));
    }
    std::sort(seen.begin(), seen.end());
    seen.erase(std::unique(seen.begin(), seen.end()), seen.end());
    for (int n : seen) if (n < n_it) needs.emplace_back(n, t);
  }
  std::sort(needs.begin(), needs.end());
  int count = 0;
  for (size_t i = 0; i < needs.size();) {
    const bool partner = i + 1 < needs.size() && needs[i + 1].first == needs[i].first;
    if (items_host) {
      if (count >= max_items) return itr::fail(ITR_ERR_INVALID, "itr_scan_plan_gt_items: more than %d items", max_items);
      int32_t* e = items_host + 4 * (size_t)count;
      e[0] = needs[i].second; e[1] = partner ? needs[i + 1].second : n_tiles; e[2] = needs[i].first; e[3] = 0;
    }
    ++count;
    i += partner ? 2 : 1;
  }
  *n_items = count;
  return ITR_OK;
}
