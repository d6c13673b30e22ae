// bzip2_enc_sort.inl -- order of all rotations of every block of a batch (replaces _blockSort, bzip2_encoder.dart:880-928).
// Included by bzip2_enc_kernels.cu inside namespace b200z::bz2e.
//
// Per block (batch-local index bl, element stride BZ2E_BSTRIDE):
//   keys (u64) / vals (u32 rotation start) / slots (u32 position in SA of the k-th unresolved element), cnt[bl] of them.
// Round 0 keys = the first 5 bytes of the rotation; later rounds = (SA slot of the element's group) << 20 | rank[pos + h].
// After a sort, equal neighbouring keys form the new groups; elements alone in their group are final and are dropped.

// first keys
__global__ void __launch_bounds__(256)
k_s_init(const uint8_t *__restrict__ blockbuf, const uint32_t *__restrict__ nblk, unsigned long long *__restrict__ keys,
         uint32_t *__restrict__ vals, uint32_t *__restrict__ slots, uint32_t *__restrict__ cnt) {
  const uint32_t bl = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const uint32_t n = nblk[bl];
  if (tile == 0 && t == 0) cnt[bl] = n;
  const uint8_t *b = blockbuf + (size_t)bl * BZ2E_BLKBYTES;
  const size_t eb = (size_t)bl * BZ2E_BSTRIDE;
  for (uint32_t j = 0; j < 8; ++j) {
    uint32_t k = tile * TS + j * 256 + t;
    if (k >= n) break;
    unsigned long long key = 0;
    uint32_t p = k;
    for (int q = 0; q < 5; ++q) {
      key = (key << 8) | b[p];
      if (++p == n) p = 0;
    }
    keys[eb + k] = key;
    vals[eb + k] = k;
    slots[eb + k] = k;
  }
}

// LSD radix pass 1/3: digit histogram of every tile, laid out [bl][bin][tile]
__global__ void __launch_bounds__(256)
k_s_hist(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ cnt, uint32_t shift,
         uint32_t *__restrict__ tile_hist) {
  __shared__ uint32_t h[256];
  const uint32_t bl = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const uint32_t c = cnt[bl], base = tile * TS;
  if (base >= c) return;
  h[t] = 0;
  __syncthreads();
  const size_t eb = (size_t)bl * BZ2E_BSTRIDE;
  for (uint32_t j = 0; j < 8; ++j) {
    uint32_t k = base + j * 256 + t;
    if (k < c) atomicAdd(&h[(uint32_t)(keys[eb + k] >> shift) & 255u], 1u);
  }
  __syncthreads();
  tile_hist[((size_t)bl * 256 + t) * NT + tile] = h[t];
}

// pass 2/3: exclusive scan in (bin, tile) order, one CTA per block
__global__ void __launch_bounds__(256)
k_s_scan(uint32_t *__restrict__ tile_hist, const uint32_t *__restrict__ cnt) {
  __shared__ uint32_t s[256];
  const uint32_t bl = blockIdx.x, t = threadIdx.x;
  const uint32_t c = cnt[bl];
  if (c == 0) return;
  const uint32_t ntl = (c + TS - 1) / TS;
  uint32_t *row = tile_hist + ((size_t)bl * 256 + t) * NT;
  uint32_t sum = 0;
  for (uint32_t k = 0; k < ntl; ++k) sum += row[k];
  s[t] = sum;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    uint32_t o = (t >= (uint32_t)d) ? s[t - d] : 0;
    __syncthreads();
    s[t] += o;
    __syncthreads();
  }
  uint32_t run = s[t] - sum;
  for (uint32_t k = 0; k < ntl; ++k) {
    uint32_t v = row[k];
    row[k] = run;
    run += v;
  }
}

// pass 3/3: stable scatter.  Element order inside a tile: warp w owns [w*256, w*256+256), 8 rounds of 32.
__global__ void __launch_bounds__(256)
k_s_scatter(const unsigned long long *__restrict__ kin, const uint32_t *__restrict__ vin, unsigned long long *__restrict__ kout,
            uint32_t *__restrict__ vout, const uint32_t *__restrict__ cnt, uint32_t shift,
            const uint32_t *__restrict__ tile_hist) {
  __shared__ uint32_t s_cnt[8][256];
  const uint32_t bl = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const uint32_t c = cnt[bl], base = tile * TS;
  if (base >= c) return;
  for (uint32_t i = t; i < 2048; i += 256) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t w = t >> 5, lane = t & 31;
  const size_t eb = (size_t)bl * BZ2E_BSTRIDE;
  unsigned long long key[8];
  uint32_t val[8], lrank[8];
  for (uint32_t r = 0; r < 8; ++r) {
    uint32_t k = base + w * 256 + r * 32 + lane;
    bool valid = k < c;
    key[r] = valid ? kin[eb + k] : 0ull;
    val[r] = valid ? vin[eb + k] : 0u;
    uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & 255u) : 256u + lane;
    uint32_t m = __match_any_sync(0xffffffffu, d);
    uint32_t leader = (uint32_t)(__ffs((int)m) - 1);
    uint32_t before = (uint32_t)__popc(m & ((1u << lane) - 1u));
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = s_cnt[w][d];
      s_cnt[w][d] = old + (uint32_t)__popc(m);
    }
    old = __shfl_sync(0xffffffffu, old, (int)leader);
    lrank[r] = old + before;
    __syncwarp();
  }
  __syncthreads();
  {
    uint32_t run = tile_hist[((size_t)bl * 256 + t) * NT + tile];
    for (uint32_t w2 = 0; w2 < 8; ++w2) {
      uint32_t v = s_cnt[w2][t];
      s_cnt[w2][t] = run;
      run += v;
    }
  }
  __syncthreads();
  for (uint32_t r = 0; r < 8; ++r) {
    uint32_t k = base + w * 256 + r * 32 + lane;
    if (k >= c) break;
    uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
    uint32_t dst = s_cnt[w][d] + lrank[r];
    kout[eb + dst] = key[r];
    vout[eb + dst] = val[r];
  }
}

// regroup 1/4: per tile, the last group start and the number of elements that stay unresolved
__global__ void __launch_bounds__(256)
k_s_groups_count(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ cnt, int *__restrict__ tile_lastflag,
                 uint32_t *__restrict__ tile_active) {
  __shared__ int s_max;
  __shared__ uint32_t s_act;
  const uint32_t bl = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const uint32_t c = cnt[bl], base = tile * TS;
  if (base >= c) return;
  if (t == 0) {
    s_max = -1;
    s_act = 0;
  }
  __syncthreads();
  const unsigned long long *kk = keys + (size_t)bl * BZ2E_BSTRIDE;
  const uint32_t k0 = base + t * 8;
  int lf = -1;
  uint32_t act = 0;
  if (k0 < c) {
    unsigned long long prev = k0 > 0 ? kk[k0 - 1] : 0ull, cur = kk[k0];
    for (uint32_t j = 0; j < 8; ++j) {
      uint32_t k = k0 + j;
      if (k >= c) break;
      unsigned long long nxt = (k + 1 < c) ? kk[k + 1] : 0ull;
      bool flag = (k == 0) || cur != prev;
      bool nflag = (k + 1 == c) || nxt != cur;
      if (flag) lf = (int)k;
      if (!(flag && nflag)) act++;
      prev = cur;
      cur = nxt;
    }
  }
  if (lf >= 0) atomicMax(&s_max, lf);
  if (act) atomicAdd(&s_act, act);
  __syncthreads();
  if (t == 0) {
    tile_lastflag[(size_t)bl * NT + tile] = s_max;
    tile_active[(size_t)bl * NT + tile] = s_act;
  }
}

// regroup 2/4: per block scans over the tiles
__global__ void __launch_bounds__(512)
k_s_block_scan(const int *__restrict__ tile_lastflag, const uint32_t *__restrict__ tile_active, const uint32_t *__restrict__ cnt,
               uint32_t *__restrict__ cnt_next, int *__restrict__ carry, uint32_t *__restrict__ act_off,
               unsigned long long *__restrict__ totals) {
  __shared__ int s_m[512];
  __shared__ uint32_t s_a[512];
  const uint32_t bl = blockIdx.x, t = threadIdx.x;
  const uint32_t c = cnt[bl];
  if (c == 0) {
    if (t == 0) cnt_next[bl] = 0;
    return;
  }
  const uint32_t ntl = (c + TS - 1) / TS;
  int lf = (t < ntl) ? tile_lastflag[(size_t)bl * NT + t] : -1;
  uint32_t ac = (t < ntl) ? tile_active[(size_t)bl * NT + t] : 0u;
  s_m[t] = lf;
  s_a[t] = ac;
  __syncthreads();
  for (int d = 1; d < 512; d <<= 1) {
    int om = (t >= (uint32_t)d) ? s_m[t - d] : -1;
    uint32_t oa = (t >= (uint32_t)d) ? s_a[t - d] : 0u;
    __syncthreads();
    if (om > s_m[t]) s_m[t] = om;
    s_a[t] += oa;
    __syncthreads();
  }
  if (t < ntl) {
    carry[(size_t)bl * NT + t] = t > 0 ? s_m[t - 1] : -1;
    act_off[(size_t)bl * NT + t] = s_a[t] - ac;
  }
  if (t == 511) {
    uint32_t total = s_a[511];
    cnt_next[bl] = total;
    if (total) {
      atomicAdd(&totals[0], (unsigned long long)total);
      atomicMax(&totals[1], (unsigned long long)total);
    }
  }
}

// regroup 3/4: SA and ranks of every element of the sorted list
__global__ void __launch_bounds__(256)
k_s_update(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ slots,
           const uint32_t *__restrict__ cnt, const int *__restrict__ carry, uint32_t *__restrict__ SA,
           uint32_t *__restrict__ rank, uint32_t *__restrict__ origptr) {
  __shared__ int s_m[256];
  const uint32_t bl = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const uint32_t c = cnt[bl], base = tile * TS;
  if (base >= c) return;
  const size_t eb = (size_t)bl * BZ2E_BSTRIDE;
  const unsigned long long *kk = keys + eb;
  const uint32_t k0 = base + t * 8;
  bool flag[8];
  int lf = -1;
  if (k0 < c) {
    unsigned long long prev = k0 > 0 ? kk[k0 - 1] : 0ull;
    for (uint32_t j = 0; j < 8; ++j) {
      uint32_t k = k0 + j;
      flag[j] = false;
      if (k >= c) continue;
      unsigned long long cur = kk[k];
      flag[j] = (k == 0) || cur != prev;
      if (flag[j]) lf = (int)k;
      prev = cur;
    }
  }
  s_m[t] = lf;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    int om = (t >= (uint32_t)d) ? s_m[t - d] : -1;
    __syncthreads();
    if (om > s_m[t]) s_m[t] = om;
    __syncthreads();
  }
  if (k0 >= c) return;
  int g = carry[(size_t)bl * NT + tile];
  if (t > 0 && s_m[t - 1] > g) g = s_m[t - 1];
  uint32_t gslot = g >= 0 ? slots[eb + (uint32_t)g] : 0u;
  for (uint32_t j = 0; j < 8; ++j) {
    uint32_t k = k0 + j;
    if (k >= c) break;
    uint32_t sl = slots[eb + k];
    if (flag[j]) gslot = sl;
    uint32_t pos = vals[eb + k];
    SA[eb + sl] = pos;
    rank[eb + pos] = gslot;
    if (pos == 0) origptr[bl] = sl;
  }
}

// regroup 4/4: compact the unresolved elements and build their next keys (h = characters already ordered)
__global__ void __launch_bounds__(256)
k_s_build(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ slots,
          const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ act_off, const uint32_t *__restrict__ rank,
          const uint32_t *__restrict__ nblk, uint32_t h, unsigned long long *__restrict__ keys2, uint32_t *__restrict__ vals2,
          uint32_t *__restrict__ slots2) {
  __shared__ uint32_t s_a[256];
  const uint32_t bl = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const uint32_t c = cnt[bl], base = tile * TS;
  if (base >= c) return;
  const size_t eb = (size_t)bl * BZ2E_BSTRIDE;
  const unsigned long long *kk = keys + eb;
  const uint32_t k0 = base + t * 8;
  bool active[8];
  uint32_t act = 0;
  if (k0 < c) {
    unsigned long long prev = k0 > 0 ? kk[k0 - 1] : 0ull, cur = kk[k0];
    for (uint32_t j = 0; j < 8; ++j) {
      uint32_t k = k0 + j;
      active[j] = false;
      if (k >= c) continue;
      unsigned long long nxt = (k + 1 < c) ? kk[k + 1] : 0ull;
      bool flag = (k == 0) || cur != prev;
      bool nflag = (k + 1 == c) || nxt != cur;
      active[j] = !(flag && nflag);
      if (active[j]) act++;
      prev = cur;
      cur = nxt;
    }
  }
  s_a[t] = act;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    uint32_t o = (t >= (uint32_t)d) ? s_a[t - d] : 0u;
    __syncthreads();
    s_a[t] += o;
    __syncthreads();
  }
  if (k0 >= c || act == 0) return;
  uint32_t dst = act_off[(size_t)bl * NT + tile] + s_a[t] - act;
  const uint32_t n = nblk[bl];
  const uint32_t hm = h % n;
  for (uint32_t j = 0; j < 8; ++j) {
    if (!active[j]) continue;
    uint32_t k = k0 + j;
    uint32_t pos = vals[eb + k];
    uint32_t p2 = pos + hm;
    if (p2 >= n) p2 -= n;
    keys2[eb + dst] = ((unsigned long long)rank[eb + pos] << 20) | rank[eb + p2];
    vals2[eb + dst] = pos;
    slots2[eb + dst] = slots[eb + k];
    dst++;
  }
}
