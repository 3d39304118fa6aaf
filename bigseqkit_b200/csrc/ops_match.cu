// ops_match.cu -- locate and grep: exact multi-pattern matching on both strands.
//
//   Locate.Before / Call (default exact path)   bigseqkit-lib/locate.go:33-204, 395-769
//   Grep.Before / grepGeneral                   bigseqkit-lib/grep.go:41-253, 367-542
//
// The reference scans every sequence once per pattern per strand with bytes.Index and
// builds a fresh reverse complement per pattern.  Here every sequence tile is staged in
// shared memory once; a rolling hash of each window is probed against a small table of
// needles (the patterns plus reverse(pair(pattern)) for the '-' strand), candidates are
// verified byte by byte.  Hits are sorted into the pinned row order: record,
// pattern as given, '+' rows ascending, '-' rows ascending on the reverse strand.
#include <algorithm>
#include <cstring>
#include <map>

#include "engine.h"
#include "op_state.h"
#include "prims.h"

namespace bsk {

static const u32 kHashB = 0x01000193u;
static const u32 kMatchTile = 4096;   // start positions per CTA
static const u32 kMatchThreads = 256;
static const u32 kPosPerThread = kMatchTile / kMatchThreads;

struct MatchArgs {
  RecViews v;
  const u32 *item_off;  // n_rec + 1: exclusive scan of tiles per record
  const u8 *nbytes;
  const u32 *nmeta, *groups, *tables;
  u32 n_groups, max_len;
  int ignore_case, circular;
  int region_on, rstart, rend;
  int mode;  // 0 locate (record hits), 1 grep (flag records)
  u64 *hitA, *hitB;
  u64 hit_cap;
  u8 *flags;
  DevStatus *st;
};

// seq.SubLocation (same rule as k_subseq_region): 0-based start + length
__device__ __forceinline__ void sub_range(u32 len, int start, int end, u32 &s0, u32 &sl) {
  s0 = 0;
  sl = 0;
  long long l = len, s = start, e = end;
  if (l == 0) return;
  if (s < 1) {
    if (s == 0) s = 1;
    else if (e < 0 && s > e) return;
    else s = (-s > l) ? 1 : l + s + 1;
  } else if (s > l) return;
  if (e > l) e = l;
  else if (e < 1) {
    if (e == 0) e = -1;
    if (-e > l) return;
    e = l + e + 1;
  }
  if (s - 1 > e) return;
  s0 = (u32)(s - 1);
  sl = (u32)(e - (s - 1));
}

__global__ void k_tiles_per_rec(RecViews v, u32 *tiles) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > v.n_rec) return;
  tiles[r] = r < v.n_rec ? (v.seq_len[r] + kMatchTile - 1) / kMatchTile : 0;
}

__global__ void __launch_bounds__(256) k_match(MatchArgs a) {
  BSK_DYN_SMEM(u8, tile);
  __shared__ u32 s_rec;
  if (threadIdx.x == 0) {
    u32 lo = 0, hi = a.v.n_rec;  // item_off[lo] <= item < item_off[hi]
    const u32 item = blockIdx.x;
    while (hi - lo > 1) {
      const u32 mid = lo + ((hi - lo) >> 1);
      if (a.item_off[mid] <= item) lo = mid;
      else hi = mid;
    }
    s_rec = lo;
  }
  __syncthreads();
  const u32 r = s_rec;
  const u32 l = a.v.seq_len[r];
  const u8 *__restrict__ s = a.v.seqb + a.v.seq_off[r];
  const u32 x0 = (blockIdx.x - a.item_off[r]) * kMatchTile;
  const u32 span = kMatchTile + a.max_len - 1;
  const u32 limit = a.circular ? 2u * l : l;  // bytes addressable from a start position
  for (u32 i = threadIdx.x; i < span; i += blockDim.x) {
    const u32 p = x0 + i;
    u8 c = 0xff;
    if (p < limit) {
      c = s[p < l ? p : p - l];
      if (a.ignore_case && c >= 'A' && c <= 'Z') c = (u8)(c + 32);
    }
    tile[i] = c;
  }
  __syncthreads();
  // search bounds per strand on the forward coordinate
  u32 lo_b[2] = {0, 0}, hi_b[2] = {l, l};
  if (a.region_on) {
    u32 s0, sl;
    sub_range(l, a.rstart, a.rend, s0, sl);
    lo_b[0] = s0;
    hi_b[0] = s0 + sl;
    lo_b[1] = l - s0 - sl;
    hi_b[1] = l - s0;
  }
  const u32 t0 = threadIdx.x * kPosPerThread;
  for (u32 g = 0; g < a.n_groups; g++) {
    const u32 L = a.groups[4 * g], tab_off = a.groups[4 * g + 3];
    const u32 tsize = a.tables[tab_off];
    const u32 *__restrict__ tab = a.tables + tab_off + 1;
    const u32 shift = 32u - (u32)(__ffs((int)tsize) - 1);
    u32 BL = 1;  // B^L
    for (u32 i = 0; i < L; i++) BL *= kHashB;
    u32 h = 0;
    for (u32 i = 0; i < L; i++) h = h * kHashB + tile[t0 + i];
    for (u32 j = 0; j < kPosPerThread; j++) {
      const u32 q = x0 + t0 + j;
      if (q < l) {
        u32 slot = (h * 0x9E3779B1u) >> shift;
        for (;;) {
          const u32 e = tab[slot];
          if (e == 0) break;
          u32 id = e - 1;
          if (a.nmeta[4 * id + 3] == h) {
            // every needle of this group with the same hash sits behind id
            const u32 gend = a.groups[4 * g + 1] + a.groups[4 * g + 2];
            for (; id < gend && a.nmeta[4 * id + 3] == h; id++) {
              const u32 ps = a.nmeta[4 * id + 2];
              const u32 strand = ps & 1u;
              bool ok = a.circular && !a.region_on ? (q + L <= 2u * l) : (q >= lo_b[strand] && q + L <= hi_b[strand]);
              if (!ok) continue;
              const u8 *nb = a.nbytes + a.nmeta[4 * id];
              for (u32 i = 0; i < L; i++)
                if (tile[t0 + j + i] != nb[i]) { ok = false; break; }
              if (!ok) continue;
              if (a.mode == 1) {
                a.flags[r] = 1;
              } else {
                const u64 idx = atomicAdd((unsigned long long *)&a.st->counters[5], 1ull);
                if (idx < a.hit_cap) {
                  u32 coord = q;
                  if (strand) coord = (q + L <= l) ? l - q - L : 2u * l - q - L;
                  a.hitA[idx] = ((u64)r << 32) | ps;
                  a.hitB[idx] = ((u64)coord << 32) | q;
                }
              }
            }
            break;
          }
          slot = (slot + 1) & (tsize - 1);
        }
      }
      h = h * kHashB + tile[t0 + j + L] - tile[t0 + j] * BL;
    }
  }
}

// -G/--non-greedy: after a hit at coordinate x the next search starts at x + m + 1 (locate.go:660)
__global__ void k_nongreedy(const u64 *__restrict__ A, const u64 *__restrict__ B, u64 n, const u32 *__restrict__ pat_meta,
                            u8 *__restrict__ keep) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i > 0 && A[i - 1] == A[i]) return;  // not a group head
  const u32 m = pat_meta[4 * ((u32)A[i] >> 1) + 3];
  u64 next_ok = 0;
  for (u64 j = i; j < n && A[j] == A[i]; j++) {
    const u64 coord = B[j] >> 32;
    if (coord >= next_ok) {
      keep[j] = 1;
      next_ok = coord + m + 1;
    } else {
      keep[j] = 0;
    }
  }
}

__device__ __forceinline__ u32 dec_digits(u64 v) {
  u32 d = 1;
  while (v >= 10) { v /= 10; d++; }
  return d;
}
__device__ __forceinline__ u8 *put_dec(u8 *o, u64 v) {
  const u32 d = dec_digits(v);
  for (u32 i = 0; i < d; i++) { o[d - 1 - i] = (u8)('0' + v % 10); v /= 10; }
  return o + d;
}
__device__ __forceinline__ u8 *put_bytes(u8 *o, const u8 *s, u32 n) {
  for (u32 i = 0; i < n; i++) o[i] = s[i];
  return o + n;
}
__device__ __forceinline__ u8 *put_lit(u8 *o, const char *s) {
  while (*s) *o++ = (u8)*s++;
  return o;
}

struct RowFmt {
  int gtf, bed, hide_matched;
};

// rows of locate.go:617-654; pass 0 computes lengths (incl. FileStore's '\n'), pass 1 writes
__global__ void k_locate_rows(const u64 *__restrict__ A, const u64 *__restrict__ B, const u8 *__restrict__ keep, u64 n,
                              RecViews v, const u32 *__restrict__ id_off, const u32 *__restrict__ id_len,
                              const u8 *__restrict__ pbytes, const u32 *__restrict__ pat_meta, RowFmt f, u32 *row_len,
                              const u64 *row_off, u8 *out, u64 out_base) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) {
    if (!out) row_len[i] = 0;
    return;
  }
  if (keep && !keep[i]) {
    if (!out) row_len[i] = 0;
    return;
  }
  const u32 r = (u32)(A[i] >> 32), ps = (u32)A[i];
  const u32 pi = ps >> 1;
  const u8 strand = (ps & 1u) ? '-' : '+';
  const u32 q = (u32)B[i];
  const u32 nlen = pat_meta[4 * pi + 1], plen = pat_meta[4 * pi + 3];
  const u64 begin = (u64)q + 1, end = (u64)q + plen;
  const u32 il = id_len[r];
  if (!out) {
    u32 n_;
    if (f.gtf) n_ = il + 17 + dec_digits(begin) + 1 + dec_digits(end) + 3 + 1 + 12 + nlen + 3;
    else if (f.bed) n_ = il + 1 + dec_digits(begin - 1) + 1 + dec_digits(end) + 1 + nlen + 3 + 1;
    else n_ = il + 1 + nlen + 1 + plen + 1 + 1 + 1 + dec_digits(begin) + 1 + dec_digits(end) + (f.hide_matched ? 0 : 1 + plen);
    row_len[i] = n_ + 1;
    return;
  }
  u8 *o = out + out_base + row_off[i];
  o = put_bytes(o, v.in + id_off[r], il);
  const u8 *name = pbytes + pat_meta[4 * pi], *pat = pbytes + pat_meta[4 * pi + 2];
  if (f.gtf) {
    o = put_lit(o, "\tSeqKit\tlocation\t");
    o = put_dec(o, begin);
    *o++ = '\t';
    o = put_dec(o, end);
    o = put_lit(o, "\t0\t");
    *o++ = strand;
    o = put_lit(o, "\t.\tgene_id \"");
    o = put_bytes(o, name, nlen);
    o = put_lit(o, "\"; ");
  } else if (f.bed) {
    *o++ = '\t';
    o = put_dec(o, begin - 1);
    *o++ = '\t';
    o = put_dec(o, end);
    *o++ = '\t';
    o = put_bytes(o, name, nlen);
    o = put_lit(o, "\t0\t");
    *o++ = strand;
  } else {
    *o++ = '\t';
    o = put_bytes(o, name, nlen);
    *o++ = '\t';
    o = put_bytes(o, pat, plen);
    *o++ = '\t';
    *o++ = strand;
    *o++ = '\t';
    o = put_dec(o, begin);
    *o++ = '\t';
    o = put_dec(o, end);
    if (!f.hide_matched) {
      *o++ = '\t';
      o = put_bytes(o, pat, plen);  // exact match: the matched text equals the (lower-cased) pattern
    }
  }
  *o++ = '\n';
}

// grep by ID / name: whole-string membership (grep.go:501-512)
__device__ __forceinline__ u64 fnv1a(const u8 *p, u32 n, int lower) {
  u64 h = 1469598103934665603ull;
  for (u32 i = 0; i < n; i++) {
    u8 c = p[i];
    if (lower && c >= 'A' && c <= 'Z') c = (u8)(c + 32);
    h = (h ^ c) * 1099511628211ull;
  }
  return h;
}
__global__ void k_grep_name(RecViews v, const u32 *__restrict__ t_off, const u32 *__restrict__ t_len, int lower,
                            const u64 *__restrict__ hashes, const u32 *__restrict__ nmeta, const u8 *__restrict__ pbytes,
                            u32 n_names, u8 *__restrict__ flags) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= v.n_rec) return;
  const u8 *t = v.in + t_off[r];
  const u32 tl = t_len[r];
  const u64 h = fnv1a(t, tl, lower);
  u32 lo = 0, hi = n_names;  // first index with hashes[idx] >= h
  while (lo < hi) {
    const u32 mid = lo + ((hi - lo) >> 1);
    if (hashes[mid] < h) lo = mid + 1;
    else hi = mid;
  }
  u8 hit = 0;
  for (u32 i = lo; i < n_names && hashes[i] == h && !hit; i++) {
    if (nmeta[2 * i + 1] != tl) continue;
    const u8 *p = pbytes + nmeta[2 * i];
    bool eq = true;
    for (u32 k2 = 0; k2 < tl; k2++) {
      u8 c = t[k2];
      if (lower && c >= 'A' && c <= 'Z') c = (u8)(c + 32);
      if (c != p[k2]) { eq = false; break; }
    }
    hit = eq ? 1 : 0;
  }
  flags[r] = hit;
}
__global__ void k_flags_to_keep(const u8 *flags, u32 n, int invert, u8 *keep, DevStatus *st) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const u8 k2 = (flags[r] != 0) != (invert != 0) ? 1 : 0;
  keep[r] = k2;
  if (k2) atomicAdd((unsigned long long *)&st->counters[6], 1ull);
}

// element offsets of the locate output: e[0] = 0 (header row), e[1 + i] = hl + row offset, e[1 + n] = total
__global__ void k_elem_fixup(u64 *e, u64 n_rows, u64 hl, u64 total) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_rows + 1) return;
  if (i == 0) e[0] = 0;
  else if (i == n_rows + 1) e[i] = total;
  else e[i] += hl;
}

// ------------------------------------------------------------------ host: needle tables
static u32 poly_hash(const std::string &s) {
  u32 h = 0;
  for (unsigned char c : s) h = h * kHashB + c;
  return h;
}
static u64 fnv1a_host(const std::string &s) {
  u64 h = 1469598103934665603ull;
  for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
  return h;
}
template <class T>
static void upload(DevBuf &b, const std::vector<T> &v, cudaStream_t s) {
  b.reserve(v.size() * sizeof(T) + 16);
  if (!v.empty()) BSK_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

int Engine::build_patterns(bool only_pos) {
  if (pats_ && pats_->alphabet == alphabet_ && pats_->only_pos == only_pos) return BSK_OK;
  delete pats_;
  pats_ = new PatternSet();
  PatternSet &ps = *pats_;
  ps.alphabet = alphabet_;
  ps.only_pos = only_pos;
  struct Needle { std::string s; u32 pat, strand, hash; };
  std::map<u32, std::vector<Needle>> by_len;
  const u8 *pair = alphabet_pair(alphabet_ == AB_NIL ? AB_UNLIMIT : alphabet_);
  for (size_t pi = 0; pi < o_.Patterns.size(); pi++) {
    const std::string &p = o_.Patterns[pi];
    by_len[(u32)p.size()].push_back(Needle{p, (u32)pi, 0, poly_hash(p)});
    if (!only_pos) {
      std::string rv(p.rbegin(), p.rend());
      for (auto &c : rv) c = (char)pair[(u8)c];
      by_len[(u32)rv.size()].push_back(Needle{rv, (u32)pi, 1, poly_hash(rv)});
    }
  }
  std::vector<u8> nbytes;
  std::vector<u32> nmeta, groups, tables;
  u32 id = 0;
  for (auto &kv : by_len) {
    auto &v = kv.second;
    std::stable_sort(v.begin(), v.end(), [](const Needle &a, const Needle &b) { return a.hash < b.hash; });
    u32 tsize = 16;
    while (tsize < v.size() * 4) tsize *= 2;
    const u32 tab_off = (u32)tables.size();
    tables.push_back(tsize);
    tables.resize(tables.size() + tsize, 0);
    u32 shift = 32;
    for (u32 t = tsize; t > 1; t >>= 1) shift--;
    groups.push_back(kv.first);
    groups.push_back(id);
    groups.push_back((u32)v.size());
    groups.push_back(tab_off);
    for (size_t i = 0; i < v.size(); i++, id++) {
      nmeta.push_back((u32)nbytes.size());
      nmeta.push_back((u32)v[i].s.size());
      nmeta.push_back((v[i].pat << 1) | v[i].strand);
      nmeta.push_back(v[i].hash);
      nbytes.insert(nbytes.end(), v[i].s.begin(), v[i].s.end());
      if (i > 0 && v[i - 1].hash == v[i].hash) continue;  // table points at the first needle of a hash run
      u32 slot = (v[i].hash * 0x9E3779B1u) >> shift;
      while (tables[tab_off + 1 + slot]) slot = (slot + 1) & (tsize - 1);
      tables[tab_off + 1 + slot] = id + 1;
    }
    ps.max_len = std::max(ps.max_len, kv.first);
  }
  ps.n_needles = id;
  ps.n_groups = (u32)by_len.size();
  for (auto &kv : by_len)
    for (auto &nd : kv.second) { ps.needle_s.push_back(nd.s); ps.needle_ps.push_back((nd.pat << 1) | nd.strand); }
  ps.kmer_built = false;
  // pattern names + text for the locate rows, hashes for grep by id/name
  std::vector<u8> pb;
  std::vector<u32> pm;
  for (size_t pi = 0; pi < o_.Patterns.size(); pi++) {
    const std::string &nm = pi < o_.PatternNames.size() ? o_.PatternNames[pi] : o_.Patterns[pi];
    pm.push_back((u32)pb.size());
    pm.push_back((u32)nm.size());
    pb.insert(pb.end(), nm.begin(), nm.end());
    pm.push_back((u32)pb.size());
    pm.push_back((u32)o_.Patterns[pi].size());
    pb.insert(pb.end(), o_.Patterns[pi].begin(), o_.Patterns[pi].end());
  }
  std::vector<std::pair<u64, u32>> hs;
  for (size_t pi = 0; pi < o_.Patterns.size(); pi++) hs.emplace_back(fnv1a_host(o_.Patterns[pi]), (u32)pi);
  std::sort(hs.begin(), hs.end());
  std::vector<u64> nh;
  std::vector<u32> nmm;
  for (auto &h : hs) {
    nh.push_back(h.first);
    nmm.push_back(pm[4 * h.second + 2]);
    nmm.push_back(pm[4 * h.second + 3]);
  }
  ps.n_names = (u32)hs.size();
  upload(ps.bytes, nbytes, stream);
  upload(ps.meta, nmeta, stream);
  upload(ps.groups, groups, stream);
  upload(ps.tables, tables, stream);
  upload(ps.pat_bytes, pb, stream);
  upload(ps.pat_meta, pm, stream);
  upload(ps.name_hash, nh, stream);
  upload(ps.name_meta, nmm, stream);
  BSK_CUDA(cudaStreamSynchronize(stream));
  if (ps.max_len > 32768) { err = "patterns longer than 32768 bytes are outside the accelerated path"; return BSK_ERR_UNSUPPORTED; }
  return BSK_OK;
}

// runs the matcher over every record of the block; mode 0 -> hits in b_op3_/b_op4_, mode 1 -> flags
int Engine::run_matcher(int mode, u8 *flags, u64 &n_hits) {
  PatternSet &ps = *pats_;
  n_hits = 0;
  const size_t R = (size_t)n_rec_ + 1;
  u32 *tiles = b_op1_.get<u32>(R);
  u32 *item_off = b_op2_.get<u32>(R);
  BSK_LAUNCH_FLAT(k_tiles_per_rec, (n_rec_ + 1 + 255) / 256, 256, 0, stream, views_, tiles);
  launches_++;
  prim::excl_scan_u32(tiles, item_off, R, b_tmp_, stream);
  u8 *hs = h_small_.as<u8>();
  prim::copy_small(hs, item_off + n_rec_, 4, stream);
  BSK_CUDA(cudaStreamSynchronize(stream));
  u32 n_items;
  memcpy(&n_items, hs, 4);
  if (n_items == 0 || ps.n_needles == 0) return BSK_OK;
  MatchArgs a;
  memset(&a, 0, sizeof a);
  a.v = views_;
  a.item_off = item_off;
  a.nbytes = ps.bytes.as<u8>();
  a.nmeta = ps.meta.as<u32>();
  a.groups = ps.groups.as<u32>();
  a.tables = ps.tables.as<u32>();
  a.n_groups = ps.n_groups;
  a.max_len = ps.max_len;
  a.ignore_case = o_.IgnoreCase;
  a.circular = o_.Circular && !(op_ == OP_GREP && o_.has_region);
  a.region_on = op_ == OP_GREP && o_.has_region;
  a.rstart = o_.region_start;
  a.rend = o_.region_end;
  a.mode = mode;
  a.flags = flags;
  a.st = d_status_;
  const size_t smem = kMatchTile + ps.max_len + 16;
  for (;;) {
    if (mode == 0) {
      if (hit_cap_ == 0) hit_cap_ = 1u << 20;
      a.hitA = b_op3_.get<u64>(hit_cap_);
      a.hitB = b_op4_.get<u64>(hit_cap_);
      a.hit_cap = hit_cap_;
      BSK_CUDA(cudaMemsetAsync(&d_status_->counters[5], 0, 8, stream));
    }
    main_begin();
    BSK_LAUNCH(k_match, n_items, kMatchThreads, smem, stream, a);
    main_end();
    launches_++;
    if (mode != 0) break;
    fetch_status();
    n_hits = h_status_->counters[5];
    if (n_hits <= hit_cap_) break;
    hit_cap_ = n_hits + n_hits / 4 + 1024;  // the buffer was too small: grow and match again
  }
  return BSK_OK;
}

// ------------------------------------------------------------------ Locate
int Engine::op_locate(BlockOut &bo, int64_t pid) {
  int rc = check_errors();
  if (rc != BSK_OK) return rc;
  bool only_pos = o_.OnlyPositiveStrand;
  if (n_rec_ && (alphabet_ == AB_UNLIMIT || alphabet_ == AB_PROTEIN)) only_pos = true;  // locate.go:424-429 (+ Q8)
  u64 n_hits = 0;
  if (n_rec_) {
    rc = build_patterns(only_pos);
    if (rc != BSK_OK) return rc;
    rc = run_matcher(0, nullptr, n_hits);
    if (rc != BSK_OK) return rc;
  }
  return locate_rows(bo, pid, n_hits);
}

// equal-length ACGT panel: 2-bit codes of the needles, first-level bitmap + fingerprint table for shared memory, exact
// table (host, once)
int Engine::build_kmer_tables() {
  PatternSet &ps = *pats_;
  if (ps.kmer_built) return BSK_OK;
  ps.kmer_built = true;
  ps.kmer_ok = false;
  if (ps.needle_s.empty() || o_.Circular) return BSK_OK;
  const size_t L = ps.needle_s[0].size();
  if (L < 2 || L > 16) return BSK_OK;
  bool upper = false, lower = false;
  for (auto &sd : ps.needle_s) {
    if (sd.size() != L) return BSK_OK;
    for (unsigned char c : sd) {
      if (c == 'A' || c == 'C' || c == 'G' || c == 'T') upper = true;
      else if (c == 'a' || c == 'c' || c == 'g' || c == 't') lower = true;
      else return BSK_OK;
    }
  }
  if (upper && lower) return BSK_OK;
  // byte classes (k_locate_tile.cu lt::C_*): valid base 8 | 16, other sequence byte 4 | 16, '\n' 0, header mark 4 | 32
  std::vector<u8> lut(256, 4 | 16);
  lut['\n'] = 0;
  lut[0x01] = 4 | 32;
  auto base = [&](char up, u8 code) {
    if (upper || o_.IgnoreCase) lut[(u8)up] = (u8)(8 | 16 | code);
    if (lower || o_.IgnoreCase) lut[(u8)(up + 32)] = (u8)(8 | 16 | code);
  };
  base('A', 0); base('C', 1); base('T', 2); base('G', 3);  // bits 1-2 of the letter, as the kernel's SWAR path packs them
  const u32 cmask = L == 16 ? 0xffffffffu : ((1u << (2 * L)) - 1u);
  struct Nd { u32 code, ps; };
  std::vector<Nd> nd;
  for (size_t i = 0; i < ps.needle_s.size(); i++) {
    u32 code = 0;
    for (unsigned char c : ps.needle_s[i]) code = code * 4u + (lut[c] & 3u);
    nd.push_back(Nd{code & cmask, ps.needle_ps[i]});
  }
  std::stable_sort(nd.begin(), nd.end(), [](const Nd &x, const Nd &y) { return x.code < y.code; });
  const u32 fbits = k::locate_tile_filter_bits();
  const u32 kmul = 0x9E3779B1u << (32 - 2 * L), kmul2 = 0x85EBCA6Bu << (32 - 2 * L);
  std::vector<u32> filter((1u << fbits) / 32, 0);
  const u32 pbits = k::locate_tile_fp_bits();
  if (nd.size() * 2 > (1u << pbits)) return BSK_OK;  // the fingerprint table would be too full: general path
  std::vector<unsigned short> fptab(1u << pbits, 0);
  u32 tsize = 16;
  while (tsize < nd.size() * 4) tsize *= 2;
  u32 tshift = 32;
  for (u32 t = tsize; t > 1; t >>= 1) tshift--;
  std::vector<u32> table(2 * (size_t)tsize, 0), codes, pss;
  for (size_t i = 0; i < nd.size(); i++) {
    codes.push_back(nd[i].code);
    pss.push_back(nd[i].ps);
    const u32 bi = (nd[i].code * kmul) >> (32 - fbits);
    filter[bi >> 5] |= 0x80000000u >> (bi & 31);  // from the top of the word: lt_probe shifts the bit into the sign
    {  // second level: 16-bit fingerprint under the second hash, linear probing (k_locate_tile.cu lt_probe2)
      const u32 h2 = nd[i].code * kmul2, want = ((h2 >> 5) & 0x7fffu) | 0x8000u;
      u32 slot = h2 >> (32 - pbits);
      while (fptab[slot] && fptab[slot] != want) slot = (slot + 1) & ((1u << pbits) - 1);
      fptab[slot] = (unsigned short)want;
    }
    if (i > 0 && nd[i - 1].code == nd[i].code) continue;  // the table points at the first needle of a code run
    u32 slot = (nd[i].code * 0x9E3779B1u) >> tshift;
    while (table[2 * slot + 1]) slot = (slot + 1) & (tsize - 1);
    table[2 * slot] = nd[i].code;
    table[2 * slot + 1] = (u32)i + 1;
  }
  ps.kL = (u32)L; ps.kfbits = fbits; ps.kmul = kmul; ps.kmul2 = kmul2; ps.kcmask = cmask; ps.ktmask = tsize - 1; ps.ktshift = tshift;
  ps.kn = (u32)nd.size();
  ps.kvmask = o_.IgnoreCase ? 0xdfdfdfdfu : 0xffffffffu;
  ps.kvbase = (lower && !o_.IgnoreCase) ? 0x61616161u : 0x41414141u;
  upload(ps.klut, lut, stream);
  upload(ps.kfilter, filter, stream);
  upload(ps.kfptab, fptab, stream);
  upload(ps.ktable, table, stream);
  upload(ps.kcode, codes, stream);
  upload(ps.kps, pss, stream);
  BSK_CUDA(cudaStreamSynchronize(stream));
  ps.kmer_ok = true;
  return BSK_OK;
}

// locate on FASTA with an equal-length ACGT panel: one streaming pass over the raw block (k_locate_tile.cu), no record
// index.  kFusedFallback when the panel / the block is outside that kernel's grammar (the general path then runs).
int Engine::op_locate_tile(const u8 *d_in, u32 n, int64_t pid, BlockOut &bo) {
  if (n == 0 || !fused_ok_ || o_.Circular || getenv("BSK_NO_LOCATE_TILE")) return kFusedFallback;
  bool fastq = false, ok = false;
  const int saved_alpha = alphabet_;
  const bool saved_known = alphabet_known_;
  auto decline = [&]() { alphabet_ = saved_alpha; alphabet_known_ = saved_known; return kFusedFallback; };
  if (!alphabet_known_ || first_block_) {
    int rc = first_record_alphabet(d_in, n, fastq, ok, true);
    if (rc != BSK_OK) return rc;
    if (!ok || fastq) return decline();
  } else if (part_fastq_) {
    return kFusedFallback;
  }
  bool only_pos = o_.OnlyPositiveStrand;
  if (alphabet_ == AB_UNLIMIT || alphabet_ == AB_PROTEIN) only_pos = true;  // locate.go:424-429 (+ Q8)
  int rc = build_patterns(only_pos);
  if (rc != BSK_OK) return decline();
  rc = build_kmer_tables();
  if (rc != BSK_OK || !pats_->kmer_ok) return decline();
  PatternSet &ps = *pats_;
  if (!n_sm_) {
    cudaDeviceProp prop;
    BSK_CUDA(cudaGetDeviceProperties(&prop, device_ >= 0 ? device_ : 0));
    n_sm_ = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
  }
  reset_status();
  BSK_CUDA(cudaEventRecord(ev_[1], stream));
  const u32 n_tiles = k::locate_tile_tiles(n);
  k::LocateTileArgs a;
  memset(&a, 0, sizeof a);
  a.in = d_in;
  a.n = n;
  a.lut = ps.klut.as<u8>();
  a.filter = ps.kfilter.as<u32>();
  a.fptab = ps.kfptab.as<unsigned short>();
  a.kmul = ps.kmul;
  a.kmul2 = ps.kmul2;
  a.L = ps.kL;
  a.cmask = ps.kcmask;
  a.vmask = ps.kvmask;
  a.vbase = ps.kvbase;
  a.table = ps.ktable.as<u32>();
  a.tmask = ps.ktmask;
  a.tshift = ps.ktshift;
  a.nd_code = ps.kcode.as<u32>();
  a.nd_ps = ps.kps.as<u32>();
  a.n_needles = ps.kn;
  a.hdr_cap = (u64)n / 32 + 1024;
  a.hdr_off = b_op1_.get<u64>((size_t)a.hdr_cap * 2);
  a.hdr_nl = a.hdr_off + a.hdr_cap;
  a.tile_nl = b_tile_cnt_.get<u32>((size_t)n_tiles + 1);
  a.st = d_status_;
  BSK_CUDA(cudaMemsetAsync(a.tile_nl + n_tiles, 0, 4, stream));
  u64 n_hits = 0, n_hdr = 0;
  for (;;) {
    if (hit_cap_ == 0) hit_cap_ = 1u << 20;
    a.hitA = b_op3_.get<u64>(hit_cap_);
    a.hitB = b_op4_.get<u64>(hit_cap_);
    a.hit_cap = hit_cap_;
    BSK_CUDA(cudaMemsetAsync(&d_status_->counters[5], 0, 16, stream));
    main_begin();
    k::locate_tile(a, n_sm_, stream);
    main_end();
    launches_++;
    fetch_status();
    if (h_status_->counters[0]) {
      main_timed_ = false;
      timings.main_launches--;
      return decline();
    }
    n_hits = h_status_->counters[5];
    n_hdr = h_status_->counters[6];
    if (n_hdr > a.hdr_cap) {  // short records: the general path indexes them
      main_timed_ = false;
      timings.main_launches--;
      return decline();
    }
    if (n_hits <= hit_cap_) break;
    hit_cap_ = n_hits + n_hits / 4 + 1024;  // the hit buffer was too small: grow and match again
    main_timed_ = false;
    timings.main_launches--;
  }
  // record table: headers in input order, newline prefix over the tiles
  n_rec_ = (u32)n_hdr;
  const size_t R = (size_t)n_rec_ + 1;
  u64 *hs_off = b_op5_.get<u64>(R * 2), *hs_nl = hs_off + R;
  prim::sort_pairs_u64_u64(a.hdr_off, hs_off, a.hdr_nl, hs_nl, n_rec_, 0, 32, b_tmp_, stream);
  u32 *tile_base = b_tile_base_.get<u32>((size_t)n_tiles + 1);
  prim::excl_scan_u32(a.tile_nl, tile_base, (size_t)n_tiles + 1, b_tmp_, stream);
  u32 *rec = b_rec_.get<u32>(R * 5);
  u32 *name_off = rec, *name_len = rec + R, *seq_start = rec + 2 * R, *seq_nl = rec + 3 * R, *seq_len = rec + 4 * R;
  k::locate_records(d_in, n, hs_off, hs_nl, tile_base, n_tiles, n_rec_, name_off, name_len, seq_start, seq_nl, seq_len, stream);
  k::locate_resolve(a.hitA, a.hitB, n_hits, hs_off, n_rec_, tile_base, seq_start, seq_nl, seq_len, ps.kL, stream);
  launches_ += 2;
  // the block state the row formatter reads
  in_ = d_in;
  n_ = n;
  fastq_ = false;
  squeezed_ = false;
  if (first_block_) part_fastq_ = false;
  memset(&ra_, 0, sizeof ra_);
  ra_.head_off = name_off;
  ra_.head_len = name_len;
  ra_.seq_len = seq_len;
  views_ = RecViews{};
  views_.in = d_in;
  views_.seqb = d_in;
  views_.qualb = d_in;
  views_.name_off = name_off;
  views_.name_len = name_len;
  views_.seq_off = seq_start;
  views_.seq_len = seq_len;
  views_.n_rec = n_rec_;
  bo.n_rec = n_rec_;
  if (n_rec_) any_record_ = true;
  timings.fused_blocks++;
  return locate_rows(bo, pid, n_hits);
}

// hits (A = record << 32 | pattern << 1 | strand, B = strand coordinate << 32 | start) in b_op3_ / b_op4_ -> rows
int Engine::locate_rows(BlockOut &bo, int64_t pid, u64 n_hits) {
  const bool tabular = !(o_.Gtf || o_.Bed);
  std::string header;
  if (tabular && pid == 0 && first_block_)  // locate.go:198-204: header row only in partition 0
    header = o_.HideMatched ? "seqID\tpatternName\tpattern\tstrand\tstart\tend\n"
                            : "seqID\tpatternName\tpattern\tstrand\tstart\tend\tmatched\n";
  const size_t H = (size_t)n_hits + 1;
  u64 *A = b_op3_.as<u64>(), *B = b_op4_.as<u64>();
  const u8 *keep = nullptr;
  u32 *row_len = b_out_len_.get<u32>(H);
  u64 *row_off = b_out_off_.get<u64>(H);
  u32 *ids = nullptr;
  RowFmt f{o_.Gtf ? 1 : 0, o_.Bed ? 1 : 0, o_.HideMatched ? 1 : 0};
  if (n_hits) {
    u64 *A2 = b_op5_.get<u64>(H), *B2 = b_op6_.get<u64>(H);
    // order: record, pattern, strand (A) then coordinate on the strand (B high word); two stable radix sorts
    prim::sort_pairs_u64_u64(B, B2, A, A2, n_hits, 32, 64, b_tmp_, stream);
    prim::sort_pairs_u64_u64(A2, A, B2, B, n_hits, 0, 64, b_tmp_, stream);
    if (o_.NonGreedy) {
      u8 *kp = b_keep_.get<u8>(H);
      BSK_LAUNCH_FLAT(k_nongreedy, (u32)((n_hits + 255) / 256), 256, 0, stream, A, B, n_hits, pats_->pat_meta.as<u32>(), kp);
      launches_++;
      keep = kp;
    }
    const size_t R = (size_t)n_rec_ + 1;
    ids = b_id_.get<u32>(R * 2);
    k::id_desc(views_, o_.IDNCBI ? 1 : 0, ids, ids + R, nullptr, nullptr, stream);
    BSK_LAUNCH_FLAT(k_locate_rows, (u32)((n_hits + 1 + 255) / 256), 256, 0, stream, A, B, keep, n_hits, views_, ids, ids + R,
                    pats_->pat_bytes.as<u8>(), pats_->pat_meta.as<u32>(), f, row_len, (const u64 *)nullptr, (u8 *)nullptr,
                    (u64)0);
    launches_ += 2;
  } else {
    BSK_CUDA(cudaMemsetAsync(row_len, 0, 4, stream));
  }
  prim::excl_scan_u32_to_u64(row_len, row_off, H, b_tmp_, stream);
  u8 *hs = h_small_.as<u8>();
  prim::copy_small(hs, row_off + n_hits, 8, stream);
  u64 n_rows = n_hits;
  u64 *elem = nullptr;
  if (keep) {
    elem = b_elem_.get<u64>(H + 2);
    prim::select_flagged_u64(row_off, keep, elem + 1, &d_status_->n_sel, n_hits, b_tmp_, stream);
    prim::copy_small(hs + 8, &d_status_->n_sel, 4, stream);
  }
  BSK_CUDA(cudaStreamSynchronize(stream));
  u64 rows_bytes;
  memcpy(&rows_bytes, hs, 8);
  if (keep) {
    u32 nsel;
    memcpy(&nsel, hs + 8, 4);
    n_rows = nsel;
  }
  const u64 hl = header.size();
  const u64 total = hl + rows_bytes;
  u8 *out = b_out_.get<u8>((size_t)total + 64);
  if (hl) {
    memcpy(hs + 64, header.data(), hl);
    BSK_CUDA(cudaMemcpyAsync(out, hs + 64, hl, cudaMemcpyHostToDevice, stream));
  }
  if (n_hits) {
    const size_t R = (size_t)n_rec_ + 1;
    BSK_LAUNCH_FLAT(k_locate_rows, (u32)((n_hits + 255) / 256), 256, 0, stream, A, B, keep, n_hits, views_, ids, ids + R,
                    pats_->pat_bytes.as<u8>(), pats_->pat_meta.as<u32>(), f, row_len, row_off, out, hl);
    launches_++;
  }
  bo.d_data = out;
  bo.n = total;
  bo.n_elem = n_rows + (hl ? 1 : 0);
  bo.d_elem_off = nullptr;
  if (want_elem_off) {
    // element offsets: [0 (header)] + hl + row offsets (+ total)
    std::vector<u64> tmp;
    u64 *e = elem ? elem : b_elem_.get<u64>(H + 2);
    if (!keep && n_hits) BSK_CUDA(cudaMemcpyAsync(e + 1, row_off, n_hits * 8, cudaMemcpyDeviceToDevice, stream));
    // shift by the header length and terminate
    BSK_LAUNCH_FLAT(k_elem_fixup, (u32)((n_rows + 2 + 255) / 256), 256, 0, stream, e, n_rows, hl, total);
    launches_++;
    bo.d_elem_off = hl ? e : e + 1;
  }
  BSK_CUDA(cudaStreamSynchronize(stream));
  return BSK_OK;
}

// grep -C: one element holding the number of matched records of the whole Call() (grep.go:526-527,538-540)
int Engine::finish_grep_count(BlockOut &bo) {
  char t[32];
  const int k2 = snprintf(t, sizeof t, "%llu\n", (unsigned long long)grep_count);
  u8 *hs = h_small_.as<u8>();
  memcpy(hs, t, (size_t)k2);
  u64 offs[2] = {0, (u64)k2};
  memcpy(hs + 64, offs, 16);
  u8 *out = b_out_.get<u8>(64);
  u64 *e = b_elem_.get<u64>(4);
  BSK_CUDA(cudaMemcpyAsync(out, hs, (size_t)k2, cudaMemcpyHostToDevice, stream));
  BSK_CUDA(cudaMemcpyAsync(e, hs + 64, 16, cudaMemcpyHostToDevice, stream));
  BSK_CUDA(cudaStreamSynchronize(stream));
  bo.d_data = out;
  bo.n = (u64)k2;
  bo.n_elem = 1;
  bo.d_elem_off = want_elem_off ? e : nullptr;
  return BSK_OK;
}

// ------------------------------------------------------------------ Grep
int Engine::op_grep(BlockOut &bo) {
  int rc = check_errors();
  if (rc != BSK_OK) return rc;
  if (first_block_) grep_count = 0;
  const size_t R = (size_t)n_rec_ + 1;
  u8 *keep = b_keep_.get<u8>(R);
  if (n_rec_) {
    bool only_pos = o_.OnlyPositiveStrand;
    if (alphabet_ == AB_UNLIMIT || alphabet_ == AB_PROTEIN) only_pos = true;  // grep.go:404-409
    if (!o_.BySeq) only_pos = true;
    rc = build_patterns(only_pos);
    if (rc != BSK_OK) return rc;
    u8 *flags = b_op7_.get<u8>(R);
    BSK_CUDA(cudaMemsetAsync(flags, 0, R, stream));
    if (o_.BySeq) {
      u64 nh = 0;
      rc = run_matcher(1, flags, nh);
      if (rc != BSK_OK) return rc;
    } else {
      const u32 *t_off = ra_.head_off, *t_len = ra_.head_len;
      if (!o_.ByName) {
        u32 *ids = b_id_.get<u32>(R * 2);
        k::id_desc(views_, o_.IDNCBI ? 1 : 0, ids, ids + R, nullptr, nullptr, stream);
        launches_++;
        t_off = ids;
        t_len = ids + R;
      }
      main_begin();
      BSK_LAUNCH_FLAT(k_grep_name, (n_rec_ + 255) / 256, 256, 0, stream, views_, t_off, t_len, o_.IgnoreCase ? 1 : 0,
                      pats_->name_hash.as<u64>(), pats_->name_meta.as<u32>(), pats_->pat_bytes.as<u8>(), pats_->n_names,
                      flags);
      main_end();
      launches_++;
    }
    BSK_CUDA(cudaMemsetAsync(&d_status_->counters[6], 0, 8, stream));
    BSK_LAUNCH_FLAT(k_flags_to_keep, (n_rec_ + 255) / 256, 256, 0, stream, flags, n_rec_, o_.InvertMatch ? 1 : 0, keep,
                    d_status_);
    launches_++;
    fetch_status();
    grep_count += h_status_->counters[6];
  }
  if (o_.Count) return BSK_OK;  // the count element is produced once per Call() by the caller
  EmitCfg cfg;
  cfg.marker = fastq_ ? '@' : '>';
  cfg.print_name = 1;
  cfg.print_seq = 1;
  cfg.print_qual = fastq_;
  cfg.plus_line = fastq_;
  cfg.reverse = 0;
  cfg.width = fastq_ ? 0 : (o_.LineWidth > 0 ? (u32)o_.LineWidth : 0);
  views_.name_off = ra_.head_off;
  views_.name_len = ra_.head_len;
  if (n_rec_ == 0) return BSK_OK;
  return emit_records(cfg, keep, nullptr, bo);
}

}  // namespace bsk
