// bzip2_enc_entropy.inl -- MTF + RUNA/RUNB (_generateMTFValues, bzip2_encoder.dart:139-265) and the Huffman stage
// (_sendMTFValues :267-745, _hbMakeCodeLengths :747-864, _hbAssignCodes :866-878, Bz2BitWriter bz2_bit_writer.dart:29-68).
// Included by bzip2_enc_kernels.cu inside namespace b200z::bz2e.

constexpr uint32_t BZ_MAX_ALPHA = 258;
constexpr uint32_t BZ_N_GROUPS = 6;
constexpr uint32_t BZ_G_SIZE = 50;
constexpr uint32_t BZ_MAX_SEL = 18002;  // 2 + 900000 / 50
constexpr uint32_t SEL_STRIDE = 18048;

struct HInfo {
  uint32_t n_groups, n_sel, alpha, hdr_bits, data_bits, n_in_use, pad0, pad1;
};
struct EncState {
  unsigned long long bitpos;
  uint32_t combined, pad;
};

// C1: last column of the sorted rotations as dense symbol numbers + last occurrence of every symbol per chunk
__global__ void __launch_bounds__(256)
k_m_lsym(const uint8_t *__restrict__ blockbuf, const uint32_t *__restrict__ SA, const uint32_t *__restrict__ nblk,
         const uint32_t *__restrict__ inuse, uint8_t *__restrict__ lsym, int *__restrict__ lastocc) {
  __shared__ uint8_t s_seq[256];
  __shared__ int s_last[256];
  const uint32_t bl = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const uint32_t n = nblk[bl], base = tile * TS;
  if (base >= n) return;
  {
    const uint32_t *u = inuse + (size_t)bl * 8;
    uint32_t cntb = 0;
    for (uint32_t w = 0; w < (t >> 5); ++w) cntb += (uint32_t)__popc(u[w]);
    cntb += (uint32_t)__popc(u[t >> 5] & ((1u << (t & 31)) - 1u));
    s_seq[t] = (uint8_t)cntb;
    s_last[t] = -1;
  }
  __syncthreads();
  const uint8_t *b = blockbuf + (size_t)bl * BZ2E_BLKBYTES;
  const size_t eb = (size_t)bl * BZ2E_BSTRIDE;
  for (uint32_t j = 0; j < 8; ++j) {
    uint32_t k = base + j * 256 + t;
    if (k >= n) break;
    uint32_t pos = SA[eb + k];
    uint32_t p = pos ? pos - 1 : n - 1;
    uint8_t sym = s_seq[b[p]];
    lsym[eb + k] = sym;
    atomicMax(&s_last[sym], (int)(k - base));
  }
  __syncthreads();
  lastocc[((size_t)bl * NT + tile) * 256 + t] = s_last[t];
}

// C2: turn the per-chunk last occurrences into "recency key at the START of the chunk":
// position of the last occurrence before the chunk, or -1-sym when the symbol has not occurred yet.
__global__ void __launch_bounds__(256)
k_m_scan_last(int *__restrict__ lastocc, const uint32_t *__restrict__ nblk) {
  const uint32_t bl = blockIdx.x, t = threadIdx.x;
  const uint32_t n = nblk[bl];
  const uint32_t ntl = (n + TS - 1) / TS;
  int run = -1 - (int)t;
  for (uint32_t tile = 0; tile < ntl; ++tile) {
    int *p = lastocc + ((size_t)bl * NT + tile) * 256 + t;
    int v = *p;
    *p = run;
    if (v >= 0) run = (int)(tile * TS) + v;
  }
}

// C3: move-to-front positions.  One THREAD per 2048-symbol chunk (the walk is serial per chunk; a warp walks 32 chunks).
// The start list of a chunk is the symbols by decreasing recency key, built cooperatively, one chunk after the other.  The
// walk keeps the first 8 list entries in registers (a BWT block's MTF positions are almost always that small): the common
// symbol costs a handful of selects and no memory access; deeper positions continue in the thread's shared-memory list.
constexpr uint32_t MTF_CPB = 64;  // chunks per CTA
__global__ void __launch_bounds__(MTF_CPB)
k_m_mtf(const uint8_t *__restrict__ lsym, const int *__restrict__ lastocc, const uint32_t *__restrict__ nblk,
        const uint32_t *__restrict__ inuse, uint8_t *__restrict__ mtfpos, int *__restrict__ tile_nzlast) {
  __shared__ int s_key[256];
  __shared__ uint8_t s_list[256][MTF_CPB];  // [position][chunk of this CTA]
  const uint32_t bl = blockIdx.y, t = threadIdx.x;
  const uint32_t n = nblk[bl];
  const uint32_t ntl = (n + TS - 1) / TS;
  const uint32_t tile0 = blockIdx.x * MTF_CPB;
  if (tile0 >= ntl) return;
  const size_t eb = (size_t)bl * BZ2E_BSTRIDE;
  uint32_t niu = 0;
  for (int w = 0; w < 8; ++w) niu += (uint32_t)__popc(inuse[(size_t)bl * 8 + w]);
  for (uint32_t q = 0; q < MTF_CPB && tile0 + q < ntl; ++q) {
    for (uint32_t i = t; i < 256; i += MTF_CPB) {
      s_key[i] = lastocc[((size_t)bl * NT + tile0 + q) * 256 + i];
      s_list[i][q] = (uint8_t)i;  // symbols >= niu never occur: they keep the tail, in index order
    }
    __syncthreads();
    for (uint32_t c = t; c < niu; c += MTF_CPB) {
      const int kc = s_key[c];
      uint32_t r = 0;
      for (uint32_t o = 0; o < niu; ++o) r += (s_key[o] > kc) ? 1u : 0u;
      s_list[r][q] = (uint8_t)c;
    }
    __syncthreads();
  }
  const uint32_t tile = tile0 + t;
  if (tile >= ntl) return;
  const uint32_t base = tile * TS, len = umin(TS, n - base);
  const uint32_t *src = reinterpret_cast<const uint32_t *>(lsym + eb + base);  // eb and base are multiples of 4
  uint32_t *dst = reinterpret_cast<uint32_t *>(mtfpos + eb + base);
  int nz = -1;
  uint32_t r0 = s_list[0][t], r1 = s_list[1][t], r2 = s_list[2][t], r3 = s_list[3][t], r4 = s_list[4][t], r5 = s_list[5][t],
           r6 = s_list[6][t], r7 = s_list[7][t];
  for (uint32_t i4 = 0; i4 < len; i4 += 4) {
    const uint32_t word = src[i4 >> 2];
    uint32_t outw = 0;
    for (uint32_t k = 0; k < 4 && i4 + k < len; ++k) {
      const uint32_t c = (word >> (8 * k)) & 0xffu;
      uint32_t j;
      if (c == r0) {
        j = 0;
      } else {
        const uint32_t m = (c == r1 ? 2u : 0u) | (c == r2 ? 4u : 0u) | (c == r3 ? 8u : 0u) | (c == r4 ? 16u : 0u) |
                           (c == r5 ? 32u : 0u) | (c == r6 ? 64u : 0u) | (c == r7 ? 128u : 0u);
        j = m ? (uint32_t)(__ffs((int)m) - 1) : 8u;
        const uint32_t last = r7;  // falls off the register window when the symbol sits deeper
        r7 = j >= 7 ? r6 : r7;
        r6 = j >= 6 ? r5 : r6;
        r5 = j >= 5 ? r4 : r5;
        r4 = j >= 4 ? r3 : r4;
        r3 = j >= 3 ? r2 : r3;
        r2 = j >= 2 ? r1 : r2;
        r1 = r0;
        r0 = c;
        if (j == 8u) {
          uint8_t prev = (uint8_t)last, cur;
          while ((cur = s_list[j][t]) != (uint8_t)c) {
            s_list[j][t] = prev;
            prev = cur;
            j++;
          }
          s_list[j][t] = prev;
        }
        nz = (int)(base + i4 + k);
      }
      outw |= j << (8 * k);
    }
    dst[i4 >> 2] = outw;
  }
  tile_nzlast[(size_t)bl * NT + tile] = nz;
}

// digits of a zero run of length r in the bijective base-2 RUNA/RUNB code
__device__ __forceinline__ uint32_t run_digits(uint32_t r) { return 31u - (uint32_t)__clz((int)(r + 1)); }

// C4: zero-run coding.  EMIT = false: symbols per tile.  EMIT = true: write mtfv + symbol frequencies.
template <bool EMIT>
__global__ void __launch_bounds__(256)
k_m_zr(const uint8_t *__restrict__ mtfpos, const int *__restrict__ tile_nzlast, const uint32_t *__restrict__ nblk,
       uint32_t *__restrict__ tile_cnt, const uint32_t *__restrict__ tile_off, const uint32_t *__restrict__ nmtf,
       const uint32_t *__restrict__ inuse, uint16_t *__restrict__ mtfv, uint32_t *__restrict__ mtf_freq) {
  __shared__ int s_m[256];
  __shared__ uint32_t s_c[256];
  __shared__ uint32_t s_f[BZ_MAX_ALPHA];
  __shared__ int s_carry;
  const uint32_t bl = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const uint32_t n = nblk[bl], base = tile * TS;
  if (base >= n) return;
  const size_t eb = (size_t)bl * BZ2E_BSTRIDE;
  const uint8_t *mp = mtfpos + eb;
  if (t == 0) {
    int c = -1;
    for (int q = (int)tile - 1; q >= 0; --q) {
      c = tile_nzlast[(size_t)bl * NT + q];
      if (c >= 0) break;
    }
    s_carry = c;
  }
  if (EMIT)
    for (uint32_t i = t; i < BZ_MAX_ALPHA; i += 256) s_f[i] = 0;
  const uint32_t k0 = base + t * 8;
  uint8_t z[9];
  int lnz = -1;
  for (uint32_t j = 0; j < 9; ++j) {
    uint32_t k = k0 + j;
    z[j] = (k < n) ? mp[k] : (uint8_t)1;  // a non-zero after the end closes the last run
    if (j < 8 && k < n && z[j]) lnz = (int)k;
  }
  s_m[t] = lnz;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    int om = (t >= (uint32_t)d) ? s_m[t - d] : -1;
    __syncthreads();
    if (om > s_m[t]) s_m[t] = om;
    __syncthreads();
  }
  int last = s_carry;
  if (t > 0 && s_m[t - 1] > last) last = s_m[t - 1];
  uint32_t cntv = 0;
  {
    int l = last;
    for (uint32_t j = 0; j < 8; ++j) {
      uint32_t k = k0 + j;
      if (k >= n) break;
      if (z[j]) {
        cntv++;
        l = (int)k;
      } else if (z[j + 1]) cntv += run_digits((uint32_t)((int)k - l));
    }
  }
  s_c[t] = cntv;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    uint32_t o = (t >= (uint32_t)d) ? s_c[t - d] : 0u;
    __syncthreads();
    s_c[t] += o;
    __syncthreads();
  }
  if (!EMIT) {
    if (t == 255) tile_cnt[(size_t)bl * NT + tile] = s_c[255];
    return;
  }
  uint32_t dst = tile_off[(size_t)bl * NT + tile] + s_c[t] - cntv;
  uint16_t *mv = mtfv + eb;
  {
    int l = last;
    for (uint32_t j = 0; j < 8; ++j) {
      uint32_t k = k0 + j;
      if (k >= n) break;
      if (z[j]) {
        mv[dst++] = (uint16_t)(z[j] + 1);
        atomicAdd(&s_f[z[j] + 1], 1u);
        l = (int)k;
      } else if (z[j + 1]) {
        uint32_t zp = (uint32_t)((int)k - l) - 1;
        for (;;) {
          uint32_t sym = zp & 1;  // RUNA = 0, RUNB = 1
          mv[dst++] = (uint16_t)sym;
          atomicAdd(&s_f[sym], 1u);
          if (zp < 2) break;
          zp = (zp - 2) / 2;
        }
      }
    }
  }
  if (tile == 0 && t == 0) {
    const uint32_t *u = inuse + (size_t)bl * 8;
    uint32_t niu = 0;
    for (int w = 0; w < 8; ++w) niu += (uint32_t)__popc(u[w]);
    uint32_t eob = niu + 1;
    mv[nmtf[bl] - 1] = (uint16_t)eob;
    atomicAdd(&s_f[eob], 1u);
  }
  __syncthreads();
  for (uint32_t i = t; i < BZ_MAX_ALPHA; i += 256)
    if (s_f[i]) atomicAdd(&mtf_freq[(size_t)bl * BZ_MAX_ALPHA + i], s_f[i]);
}

// per block: exclusive scan of the tile symbol counts; nmtf = total + 1 (EOB)
__global__ void __launch_bounds__(512)
k_m_zr_scan(const uint32_t *__restrict__ tile_cnt, const uint32_t *__restrict__ nblk, uint32_t *__restrict__ tile_off,
            uint32_t *__restrict__ nmtf) {
  __shared__ uint32_t s[512];
  const uint32_t bl = blockIdx.x, t = threadIdx.x;
  const uint32_t n = nblk[bl];
  const uint32_t ntl = (n + TS - 1) / TS;
  uint32_t v = (t < ntl) ? tile_cnt[(size_t)bl * NT + t] : 0u;
  s[t] = v;
  __syncthreads();
  for (int d = 1; d < 512; d <<= 1) {
    uint32_t o = (t >= (uint32_t)d) ? s[t - d] : 0u;
    __syncthreads();
    s[t] += o;
    __syncthreads();
  }
  if (t < ntl) tile_off[(size_t)bl * NT + t] = s[t] - v;
  if (t == 511) nmtf[bl] = s[511] + 1;
}

// ---------------------------------------------------------------------------------------------
// Huffman stage
// ---------------------------------------------------------------------------------------------
// _hbMakeCodeLengths :747-864 (one thread)
__device__ void hb_make_code_lengths(uint8_t *len, const int *freq, int alpha, int max_len, int *heap, int *weight, int *parent) {
  for (int i = 0; i < alpha; i++) weight[i + 1] = (freq[i] == 0 ? 1 : freq[i]) << 8;
  for (;;) {
    int n_nodes = alpha, n_heap = 0;
    heap[0] = 0;
    weight[0] = 0;
    parent[0] = -2;
    for (int i = 1; i <= alpha; i++) {
      parent[i] = -1;
      n_heap++;
      heap[n_heap] = i;
      int zz = n_heap, tmp = heap[zz];
      while (weight[tmp] < weight[heap[zz >> 1]]) {
        heap[zz] = heap[zz >> 1];
        zz >>= 1;
      }
      heap[zz] = tmp;
    }
    while (n_heap > 1) {
      int n1, n2;
      for (int rep = 0; rep < 2; ++rep) {
        int top = heap[1];
        heap[1] = heap[n_heap];
        n_heap--;
        int zz = 1, tmp = heap[zz];
        for (;;) {
          int yy = zz << 1;
          if (yy > n_heap) break;
          if (yy < n_heap && weight[heap[yy + 1]] < weight[heap[yy]]) yy++;
          if (weight[tmp] < weight[heap[yy]]) break;
          heap[zz] = heap[yy];
          zz = yy;
        }
        heap[zz] = tmp;
        if (rep == 0) n1 = top;
        else n2 = top;
      }
      n_nodes++;
      parent[n1] = parent[n2] = n_nodes;
      {
        uint32_t w1 = (uint32_t)weight[n1], w2 = (uint32_t)weight[n2];
        uint32_t d1 = w1 & 0xffu, d2 = w2 & 0xffu;
        weight[n_nodes] = (int)(((w1 & 0xffffff00u) + (w2 & 0xffffff00u)) | (1u + (d1 > d2 ? d1 : d2)));
      }
      parent[n_nodes] = -1;
      n_heap++;
      heap[n_heap] = n_nodes;
      int zz = n_heap, tmp = heap[zz];
      while (weight[tmp] < weight[heap[zz >> 1]]) {
        heap[zz] = heap[zz >> 1];
        zz >>= 1;
      }
      heap[zz] = tmp;
    }
    bool too_long = false;
    for (int i = 1; i <= alpha; i++) {
      int j = 0, k = i;
      while (parent[k] >= 0) {
        k = parent[k];
        j++;
      }
      len[i - 1] = (uint8_t)j;
      if (j > max_len) too_long = true;
    }
    if (!too_long) break;
    for (int i = 1; i <= alpha; i++) {
      int j = weight[i] >> 8;
      j = 1 + (j / 2);
      weight[i] = j << 8;
    }
  }
}

// D1: coding tables of one block (one CTA)
__global__ void __launch_bounds__(512)
k_h_tables(const uint16_t *__restrict__ mtfv, const uint32_t *__restrict__ nmtf_a, const uint32_t *__restrict__ mtf_freq,
           const uint32_t *__restrict__ inuse, uint8_t *__restrict__ selector, uint8_t *__restrict__ selmtf,
           uint8_t *__restrict__ lens, uint32_t *__restrict__ codes, HInfo *__restrict__ hinfo,
           uint32_t *__restrict__ tile_bitoff) {
  __shared__ uint8_t s_len[BZ_N_GROUPS][BZ_MAX_ALPHA + 2];
  __shared__ int s_rfreq[BZ_N_GROUPS][BZ_MAX_ALPHA];
  __shared__ int s_heap[BZ_N_GROUPS][BZ_MAX_ALPHA + 2];
  __shared__ int s_weight[BZ_N_GROUPS][BZ_MAX_ALPHA * 2];
  __shared__ int s_parent[BZ_N_GROUPS][BZ_MAX_ALPHA * 2];
  __shared__ uint32_t s_tile[NET + 1];
  __shared__ uint32_t s_misc[8];
  const uint32_t bl = blockIdx.x, t = threadIdx.x;
  const uint32_t nmtf = nmtf_a[bl];
  const uint16_t *mv = mtfv + (size_t)bl * BZ2E_BSTRIDE;
  const uint32_t *mf = mtf_freq + (size_t)bl * BZ_MAX_ALPHA;
  uint8_t *sel = selector + (size_t)bl * SEL_STRIDE;
  uint8_t *smtf = selmtf + (size_t)bl * SEL_STRIDE;
  uint32_t niu = 0;
  for (int w = 0; w < 8; ++w) niu += (uint32_t)__popc(inuse[(size_t)bl * 8 + w]);
  const uint32_t alpha = niu + 2;
  uint32_t n_groups;
  if (nmtf < 200) n_groups = 2;
  else if (nmtf < 600) n_groups = 3;
  else if (nmtf < 1200) n_groups = 4;
  else if (nmtf < 2400) n_groups = 5;
  else n_groups = 6;
  const uint32_t n_sel = (nmtf + BZ_G_SIZE - 1) / BZ_G_SIZE;
  if (t == 0) {  // initial partition :300-340
    int n_part = (int)n_groups, rem_f = (int)nmtf, gs = 0;
    while (n_part > 0) {
      int t_freq = rem_f / n_part, ge = gs - 1, a_freq = 0;
      while (a_freq < t_freq && ge < (int)alpha - 1) {
        ge++;
        a_freq += (int)mf[ge];
      }
      if (ge > gs && n_part != (int)n_groups && n_part != 1 && (((int)n_groups - n_part) % 2 == 1)) {
        a_freq -= (int)mf[ge];
        ge--;
      }
      for (int v = 0; v < (int)alpha; v++) s_len[n_part - 1][v] = (v >= gs && v <= ge) ? 0 : 15;
      n_part--;
      gs = ge + 1;
      rem_f -= a_freq;
    }
  }
  __syncthreads();
  for (int iter = 0; iter < 4; ++iter) {
    for (uint32_t i = t; i < BZ_N_GROUPS * BZ_MAX_ALPHA; i += 512) (&s_rfreq[0][0])[i] = 0;
    __syncthreads();
    for (uint32_t g = t; g < n_sel; g += 512) {
      const uint32_t gs = g * BZ_G_SIZE, ge = umin(gs + BZ_G_SIZE, nmtf);
      uint32_t cost[BZ_N_GROUPS] = {0, 0, 0, 0, 0, 0};
      for (uint32_t i = gs; i < ge; ++i) {
        uint32_t icv = mv[i];
        for (uint32_t q = 0; q < BZ_N_GROUPS; ++q) cost[q] += s_len[q][icv];
      }
      uint32_t bc = 999999999u, bt = 0;
      for (uint32_t q = 0; q < n_groups; ++q) {
        uint32_t cq = cost[q] & 0xffffu;  // cost is a Uint16List (:268)
        if (cq < bc) {
          bc = cq;
          bt = q;
        }
      }
      sel[g] = (uint8_t)bt;
      // RUNA / RUNB are a large part of a BWT block: count them in registers, one shared-memory atomic per group
      uint32_t c0 = 0, c1 = 0;
      for (uint32_t i = gs; i < ge; ++i) {
        const uint32_t icv = mv[i];
        c0 += icv == 0u;
        c1 += icv == 1u;
        if (icv > 1u) atomicAdd(&s_rfreq[bt][icv], 1);
      }
      if (c0) atomicAdd(&s_rfreq[bt][0], (int)c0);
      if (c1) atomicAdd(&s_rfreq[bt][1], (int)c1);
    }
    __syncthreads();
    if (t < n_groups) hb_make_code_lengths(s_len[t], s_rfreq[t], (int)alpha, 17, s_heap[t], s_weight[t], s_parent[t]);
    __syncthreads();
  }
  // selector MTF :640-657.  The list over <= 6 tables at any selector is the tables by recency (never-used ones keep
  // their initial order at the end), so every thread rebuilds it for the start of its own run of selectors by looking
  // back, then walks the run; selectors are staged in shared memory (the tree arrays are free by now).
  uint8_t *s_sel = reinterpret_cast<uint8_t *>(&s_weight[0][0]);  // 6 * 516 * 4 = 12 384 B
  uint8_t *s_sel2 = reinterpret_cast<uint8_t *>(&s_parent[0][0]); // the next 12 384 B
  if (t == 0) s_misc[0] = 0;
  for (uint32_t i = t; i < n_sel; i += 512) {
    const uint8_t v = sel[i];
    if (i < 12384u) s_sel[i] = v;
    else s_sel2[i - 12384u] = v;
  }
  __syncthreads();
  {
    const uint32_t per = (n_sel + 511) / 512;
    const uint32_t lo = umin(n_sel, t * per), hi = umin(n_sel, lo + per);
    if (hi > lo) {
      uint8_t pos[BZ_N_GROUPS];
      uint32_t np_ = 0, seen = 0;
      for (int i = (int)lo - 1; i >= 0 && np_ < n_groups; --i) {
        const uint32_t v = (uint32_t)i < 12384u ? s_sel[i] : s_sel2[i - 12384];
        if (!((seen >> v) & 1u)) {
          seen |= 1u << v;
          pos[np_++] = (uint8_t)v;
        }
      }
      for (uint32_t v = 0; v < n_groups; ++v)
        if (!((seen >> v) & 1u)) pos[np_++] = (uint8_t)v;
      uint32_t bits = 0;
      for (uint32_t i = lo; i < hi; ++i) {
        const uint8_t ll = i < 12384u ? s_sel[i] : s_sel2[i - 12384u];
        uint32_t j = 0;
        uint8_t tmp = pos[0];
        while (ll != tmp) {
          j++;
          const uint8_t tmp2 = tmp;
          tmp = pos[j];
          pos[j] = tmp2;
        }
        pos[0] = tmp;
        smtf[i] = (uint8_t)j;
        bits += j + 1;
      }
      atomicAdd(&s_misc[0], bits);
    }
  }
  if (t >= 32 && t < 32 + n_groups) {
    const uint32_t q = t - 32;
    uint32_t min_len = 32, max_len = 0;
    for (uint32_t i = 0; i < alpha; i++) {
      uint32_t l = s_len[q][i];
      if (l > max_len) max_len = l;
      if (l < min_len) min_len = l;
    }
    uint32_t vec = 0;
    uint32_t *cd = codes + ((size_t)bl * BZ_N_GROUPS + q) * BZ_MAX_ALPHA;
    for (uint32_t nn = min_len; nn <= max_len; nn++) {
      for (uint32_t i = 0; i < alpha; i++)
        if (s_len[q][i] == nn) cd[i] = vec++;
      vec <<= 1;
    }
    uint32_t bits = 5;
    int curr = s_len[q][0];
    for (uint32_t i = 0; i < alpha; i++) {
      int l = s_len[q][i];
      bits += 1 + 2 * (uint32_t)(l > curr ? l - curr : curr - l);
      curr = l;
    }
    s_misc[1 + q] = bits;
    uint8_t *lo = lens + ((size_t)bl * BZ_N_GROUPS + q) * BZ_MAX_ALPHA;
    for (uint32_t i = 0; i < alpha; i++) lo[i] = s_len[q][i];
  }
  for (uint32_t i = t; i <= NET; i += 512) s_tile[i] = 0;
  __syncthreads();
  // bit cost of every group under the final tables, accumulated per emission tile (40 groups)
  for (uint32_t g = t; g < n_sel; g += 512) {
    const uint32_t gs = g * BZ_G_SIZE, ge = umin(gs + BZ_G_SIZE, nmtf);
    const uint32_t q = sel[g];
    uint32_t bits = 0;
    for (uint32_t i = gs; i < ge; ++i) bits += s_len[q][mv[i]];
    atomicAdd(&s_tile[g / (ET / BZ_G_SIZE)], bits);
  }
  __syncthreads();
  if (t == 0) {
    uint32_t run = 0;
    const uint32_t ntile = (nmtf + ET - 1) / ET;
    uint32_t *to = tile_bitoff + (size_t)bl * (NET + 1);
    for (uint32_t i = 0; i < ntile; ++i) {
      to[i] = run;
      run += s_tile[i];
    }
    uint32_t used16 = 0;
    for (uint32_t i = 0; i < 16; i++) {
      uint32_t w = inuse[(size_t)bl * 8 + (i >> 1)];
      if ((w >> ((i & 1) * 16)) & 0xffffu) used16++;
    }
    uint32_t hdr = 48 + 32 + 1 + 24 + 16 + 16 * used16 + 3 + 15 + s_misc[0];
    for (uint32_t q = 0; q < n_groups; ++q) hdr += s_misc[1 + q];
    HInfo hi;
    hi.n_groups = n_groups;
    hi.n_sel = n_sel;
    hi.alpha = alpha;
    hi.hdr_bits = hdr;
    hi.data_bits = run;
    hi.n_in_use = niu;
    hi.pad0 = hi.pad1 = 0;
    hinfo[bl] = hi;
  }
}

// MSB-first bit placement into a zeroed buffer of big-endian 32-bit words
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ void put_bits(uint32_t *__restrict__ out, unsigned long long bitpos, uint32_t nbits, uint32_t value) {
  if (nbits == 0) return;
  const unsigned long long w = bitpos >> 5;
  const uint32_t o = (uint32_t)bitpos & 31u;
  const unsigned long long v64 = (unsigned long long)value << (64 - o - nbits);
  const uint32_t hi = (uint32_t)(v64 >> 32), lo = (uint32_t)v64;
  if (hi) atomicOr(&out[w], bswap32(hi));
  if (lo) atomicOr(&out[w + 1], bswap32(lo));
}

// D2: bit offset of every block of the batch, running combined CRC (encodeStream :66-67), stream header
__global__ void k_h_offsets(const HInfo *__restrict__ hinfo, const uint32_t *__restrict__ block_crc, uint32_t nb, uint32_t first,
                            EncState *__restrict__ st, unsigned long long *__restrict__ bit_off, uint32_t *__restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long pos = st->bitpos;
  uint32_t comb = st->combined;
  if (first) {
    pos = 32;
    comb = 0;
    (void)out;
  }
  for (uint32_t b = 0; b < nb; ++b) {
    bit_off[b] = pos;
    pos += (unsigned long long)hinfo[b].hdr_bits + hinfo[b].data_bits;
    comb = ((comb << 1) | (comb >> 31)) ^ block_crc[b];
  }
  st->bitpos = pos;
  st->combined = comb;
}

// D3: block header, symbol map, selectors and the coding tables (:659-722)
__global__ void __launch_bounds__(256)
k_h_emit_header(const HInfo *__restrict__ hinfo, const unsigned long long *__restrict__ bit_off,
                const uint32_t *__restrict__ block_crc, const uint32_t *__restrict__ origptr, const uint32_t *__restrict__ inuse,
                const uint8_t *__restrict__ selmtf, const uint8_t *__restrict__ lens, uint32_t *__restrict__ out) {
  __shared__ uint32_t s[256];
  __shared__ unsigned long long s_pos;
  const uint32_t bl = blockIdx.x, t = threadIdx.x;
  const HInfo hi = hinfo[bl];
  const uint8_t *smtf = selmtf + (size_t)bl * SEL_STRIDE;
  if (t == 0) {
    unsigned long long p = bit_off[bl];
    put_bits(out, p, 24, 0x314159u);
    p += 24;
    put_bits(out, p, 24, 0x265359u);
    p += 24;
    put_bits(out, p, 32, block_crc[bl]);
    p += 32;
    put_bits(out, p, 1, 0);
    p += 1;
    put_bits(out, p, 24, origptr[bl]);
    p += 24;
    uint32_t m16 = 0;
    for (uint32_t i = 0; i < 16; i++) {
      uint32_t w = (inuse[(size_t)bl * 8 + (i >> 1)] >> ((i & 1) * 16)) & 0xffffu;
      if (w) m16 |= 1u << (15 - i);
    }
    put_bits(out, p, 16, m16);
    p += 16;
    for (uint32_t i = 0; i < 16; i++) {
      uint32_t w = (inuse[(size_t)bl * 8 + (i >> 1)] >> ((i & 1) * 16)) & 0xffffu;
      if (!w) continue;
      put_bits(out, p, 16, __brev(w) >> 16);  // symbol i*16 + j is written j-th
      p += 16;
    }
    put_bits(out, p, 3, hi.n_groups);
    p += 3;
    put_bits(out, p, 15, hi.n_sel);
    p += 15;
    s_pos = p;
  }
  // selectors: j ones and a zero each
  const uint32_t per = (hi.n_sel + 255) / 256;
  const uint32_t lo = umin(hi.n_sel, t * per), hi_i = umin(hi.n_sel, lo + per);
  uint32_t bits = 0;
  for (uint32_t i = lo; i < hi_i; ++i) bits += smtf[i] + 1u;
  s[t] = bits;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    uint32_t o = (t >= (uint32_t)d) ? s[t - d] : 0u;
    __syncthreads();
    s[t] += o;
    __syncthreads();
  }
  {
    unsigned long long p = s_pos + (s[t] - bits);
    for (uint32_t i = lo; i < hi_i; ++i) {
      uint32_t j = smtf[i];
      put_bits(out, p, j + 1, ((1u << j) - 1u) << 1);
      p += j + 1;
    }
  }
  // coding tables, one thread each
  if (t < hi.n_groups) {
    unsigned long long p = s_pos + s[255];
    for (uint32_t q = 0; q <= t; ++q) {
      const uint8_t *l = lens + ((size_t)bl * BZ_N_GROUPS + q) * BZ_MAX_ALPHA;
      int curr = l[0];
      if (q == t) put_bits(out, p, 5, (uint32_t)curr);
      p += 5;
      for (uint32_t i = 0; i < hi.alpha; i++) {
        int li = l[i];
        while (curr < li) {
          if (q == t) put_bits(out, p, 2, 2);
          p += 2;
          curr++;
        }
        while (curr > li) {
          if (q == t) put_bits(out, p, 2, 3);
          p += 2;
          curr--;
        }
        if (q == t) put_bits(out, p, 1, 0);
        p += 1;
      }
    }
  }
}

// D4: the coded symbols, one tile of 2000 symbols (40 groups) per CTA, 8 symbols per thread
__global__ void __launch_bounds__(256)
k_h_emit_data(const uint16_t *__restrict__ mtfv, const uint32_t *__restrict__ nmtf_a, const HInfo *__restrict__ hinfo,
              const unsigned long long *__restrict__ bit_off, const uint32_t *__restrict__ tile_bitoff,
              const uint8_t *__restrict__ selector, const uint8_t *__restrict__ lens, const uint32_t *__restrict__ codes,
              uint32_t *__restrict__ out) {
  __shared__ uint32_t s[256];
  const uint32_t bl = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const uint32_t nmtf = nmtf_a[bl];
  const uint32_t base = tile * ET;
  if (base >= nmtf) return;
  const uint16_t *mv = mtfv + (size_t)bl * BZ2E_BSTRIDE;
  const uint8_t *sel = selector + (size_t)bl * SEL_STRIDE;
  const uint8_t *lb = lens + (size_t)bl * BZ_N_GROUPS * BZ_MAX_ALPHA;
  const uint32_t *cb = codes + (size_t)bl * BZ_N_GROUPS * BZ_MAX_ALPHA;
  const uint32_t i0 = base + t * 8;
  uint32_t l[8], c[8], tot = 0;
  for (uint32_t j = 0; j < 8; ++j) {
    uint32_t i = i0 + j;
    l[j] = 0;
    c[j] = 0;
    if (t * 8 + j < ET && i < nmtf) {
      uint32_t q = sel[i / BZ_G_SIZE], sym = mv[i];
      l[j] = lb[q * BZ_MAX_ALPHA + sym];
      c[j] = cb[q * BZ_MAX_ALPHA + sym];
      tot += l[j];
    }
  }
  s[t] = tot;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    uint32_t o = (t >= (uint32_t)d) ? s[t - d] : 0u;
    __syncthreads();
    s[t] += o;
    __syncthreads();
  }
  if (tot == 0) return;
  const HInfo hi = hinfo[bl];
  unsigned long long p = bit_off[bl] + hi.hdr_bits + tile_bitoff[(size_t)bl * (NET + 1) + tile] + (s[t] - tot);
  unsigned long long w = p >> 5;
  uint32_t fill = (uint32_t)p & 31u, cur = 0;
  for (uint32_t j = 0; j < 8; ++j) {
    uint32_t rem = l[j];
    const uint32_t code = c[j];
    while (rem) {
      uint32_t take = umin(32u - fill, rem);
      uint32_t piece = (code >> (rem - take)) & (take == 32 ? 0xffffffffu : ((1u << take) - 1u));
      cur |= piece << (32 - fill - take);
      fill += take;
      rem -= take;
      if (fill == 32) {
        atomicOr(&out[w], bswap32(cur));
        w++;
        cur = 0;
        fill = 0;
      }
    }
  }
  if (fill && cur) atomicOr(&out[w], bswap32(cur));
}

// end of stream: magic, combined CRC, pad to a byte (:70-77)
__global__ void k_h_finish(EncState *__restrict__ st, uint32_t *__restrict__ out, unsigned long long *__restrict__ out_bytes) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long p = st->bitpos;
  put_bits(out, 0, 32, 0x425a6839u);  // "BZh9"
  put_bits(out, p, 24, 0x177245u);
  p += 24;
  put_bits(out, p, 24, 0x385090u);
  p += 24;
  put_bits(out, p, 32, st->combined);
  p += 32;
  st->bitpos = p;
  out_bytes[0] = (p + 7) >> 3;
}
