// bzip2_enc_serial.inl -- rotation order of PERIODIC blocks (blocks with identical rotations).
// Included by bzip2_enc_kernels.cu inside namespace b200z::bz2e.
//
// For such blocks the order among the identical rotations, and with it origPtr, depends on the exact sequence of swaps
// of the reference's sort (_mainSort with its work budget, bzip2_encoder.dart:1247-2011, falling back to _fallbackSort
// :930-1245 when the budget runs out or nblock < 10000, _blockSort :880-928).  Nothing about that order is canonical,
// so these (rare) blocks run the reference's algorithm as it is, one device thread per block.  All other blocks never
// come here: their order is unique and the batched prefix-doubling sort produces it.

namespace serial {
constexpr int N_RADIX = 2, N_QSORT = 12, N_SHELL = 18, N_OVERSHOOT = N_RADIX + N_QSORT + N_SHELL + 2;

__device__ void fb_simple_sort(uint32_t *fmap, const uint32_t *eclass, int lo, int hi) {
  if (lo == hi) return;
  if (hi - lo > 3) {
    for (int i = hi - 4; i >= lo; i--) {
      int tmp = (int)fmap[i];
      uint32_t ec = eclass[tmp];
      int j;
      for (j = i + 4; j <= hi && ec > eclass[fmap[j]]; j += 4) fmap[j - 4] = fmap[j];
      fmap[j - 4] = (uint32_t)tmp;
    }
  }
  for (int i = hi - 1; i >= lo; i--) {
    int tmp = (int)fmap[i];
    uint32_t ec = eclass[tmp];
    int j;
    for (j = i + 1; j <= hi && ec > eclass[fmap[j]]; j++) fmap[j - 1] = fmap[j];
    fmap[j - 1] = (uint32_t)tmp;
  }
}
__device__ __forceinline__ void swp(uint32_t &a, uint32_t &b) {
  uint32_t t = a;
  a = b;
  b = t;
}
__device__ bool fb_qsort3(uint32_t *fmap, const uint32_t *eclass, int lo_st, int hi_st) {
  int stack_lo[100], stack_hi[100];
  int sp = 0;
  uint32_t r = 0;
  stack_lo[sp] = lo_st;
  stack_hi[sp] = hi_st;
  sp++;
  while (sp > 0) {
    if (sp >= 99) return false;
    sp--;
    int lo = stack_lo[sp], hi = stack_hi[sp];
    if (hi - lo < 10) {
      fb_simple_sort(fmap, eclass, lo, hi);
      continue;
    }
    r = ((r * 7621) + 1) % 32768;
    uint32_t r3 = r % 3;
    long long med;
    if (r3 == 0) med = eclass[fmap[lo]];
    else if (r3 == 1) med = eclass[fmap[(lo + hi) >> 1]];
    else med = eclass[fmap[hi]];
    int un_lo = lo, lt_lo = lo, un_hi = hi, gt_hi = hi;
    for (;;) {
      for (;;) {
        if (un_lo > un_hi) break;
        long long nn = (long long)eclass[fmap[un_lo]] - med;
        if (nn == 0) {
          swp(fmap[un_lo], fmap[lt_lo]);
          lt_lo++;
          un_lo++;
          continue;
        }
        if (nn > 0) break;
        un_lo++;
      }
      for (;;) {
        if (un_lo > un_hi) break;
        long long nn = (long long)eclass[fmap[un_hi]] - med;
        if (nn == 0) {
          swp(fmap[un_hi], fmap[gt_hi]);
          gt_hi--;
          un_hi--;
          continue;
        }
        if (nn < 0) break;
        un_hi--;
      }
      if (un_lo > un_hi) break;
      swp(fmap[un_lo], fmap[un_hi]);
      un_lo++;
      un_hi--;
    }
    if (un_hi != un_lo - 1) return false;
    if (gt_hi < lt_lo) continue;
    int nn = (lt_lo - lo) < (un_lo - lt_lo) ? (lt_lo - lo) : (un_lo - lt_lo);
    for (int a = lo, b = un_lo - nn, c = nn; c > 0; a++, b++, c--) swp(fmap[a], fmap[b]);
    int mm = (hi - gt_hi) < (gt_hi - un_hi) ? (hi - gt_hi) : (gt_hi - un_hi);
    for (int a = un_lo, b = hi - mm + 1, c = mm; c > 0; a++, b++, c--) swp(fmap[a], fmap[b]);
    nn = lo + un_lo - lt_lo - 1;
    mm = hi - (gt_hi - un_hi) + 1;
    if (nn - lo > hi - mm) {
      stack_lo[sp] = lo;
      stack_hi[sp] = nn;
      sp++;
      stack_lo[sp] = mm;
      stack_hi[sp] = hi;
      sp++;
    } else {
      stack_lo[sp] = mm;
      stack_hi[sp] = hi;
      sp++;
      stack_lo[sp] = lo;
      stack_hi[sp] = nn;
      sp++;
    }
  }
  return true;
}

// _fallbackSort.  block: the nblock bytes; fmap: result; eclass, bhtab: scratch.
__device__ bool fallback_sort(uint32_t *fmap, uint32_t *eclass, uint32_t *bhtab, const uint8_t *block, int nblock) {
  int ftab[257];
  for (int i = 0; i < 257; i++) ftab[i] = 0;
  for (int i = 0; i < nblock; i++) ftab[block[i]]++;
  for (int i = 1; i < 257; i++) ftab[i] += ftab[i - 1];
  for (int i = 0; i < nblock; i++) {
    int j = block[i];
    int k = ftab[j] - 1;
    ftab[j] = k;
    fmap[k] = (uint32_t)i;
  }
  const int n_bhtab = 2 + (nblock / 32);
  for (int i = 0; i < n_bhtab; i++) bhtab[i] = 0;
#define BZ_SET_BH(zz) bhtab[(zz) >> 5] |= (1u << ((zz)&31))
#define BZ_CLEAR_BH(zz) bhtab[(zz) >> 5] &= ~(1u << ((zz)&31))
#define BZ_ISSET_BH(zz) (bhtab[(zz) >> 5] & (1u << ((zz)&31)))
#define BZ_WORD_BH(zz) bhtab[(zz) >> 5]
#define BZ_UNALIGNED_BH(zz) ((zz)&0x01f)
  for (int i = 0; i < 256; i++) BZ_SET_BH(ftab[i]);
  for (int i = 0; i < 32; i++) {
    BZ_SET_BH(nblock + 2 * i);
    BZ_CLEAR_BH(nblock + 2 * i + 1);
  }
  int H = 1;
  for (;;) {
    int j = 0;
    for (int i = 0; i < nblock; i++) {
      if (BZ_ISSET_BH(i)) j = i;
      int k = (int)fmap[i] - H;
      if (k < 0) k += nblock;
      eclass[k] = (uint32_t)j;
    }
    int n_not_done = 0, r = -1;
    for (;;) {
      int k = r + 1;
      while (BZ_ISSET_BH(k) && BZ_UNALIGNED_BH(k)) k++;
      if (BZ_ISSET_BH(k)) {
        while (BZ_WORD_BH(k) == 0xffffffffu) k += 32;
        while (BZ_ISSET_BH(k)) k++;
      }
      int l = k - 1;
      if (l >= nblock) break;
      while (!BZ_ISSET_BH(k) && BZ_UNALIGNED_BH(k)) k++;
      if (!BZ_ISSET_BH(k)) {
        while (BZ_WORD_BH(k) == 0x00000000u) k += 32;
        while (!BZ_ISSET_BH(k)) k++;
      }
      r = k - 1;
      if (r >= nblock) break;
      if (r > l) {
        n_not_done += (r - l + 1);
        if (!fb_qsort3(fmap, eclass, l, r)) return false;
        int cc = -1;
        for (int i = l; i <= r; i++) {
          int cc1 = (int)eclass[fmap[i]];
          if (cc != cc1) {
            BZ_SET_BH(i);
            cc = cc1;
          }
        }
      }
    }
    H *= 2;
    if (H > nblock || n_not_done == 0) break;
  }
#undef BZ_SET_BH
#undef BZ_CLEAR_BH
#undef BZ_ISSET_BH
#undef BZ_WORD_BH
#undef BZ_UNALIGNED_BH
  return true;
}

__device__ bool main_gtu(uint32_t i1, uint32_t i2, const uint8_t *block, const uint16_t *quadrant, uint32_t nblock,
                         int *budget) {
  if (i1 == i2) return false;
  for (int t = 0; t < 12; ++t) {
    uint8_t c1 = block[i1], c2 = block[i2];
    if (c1 != c2) return c1 > c2;
    i1++;
    i2++;
  }
  int k = (int)nblock + 8;
  do {
    for (int t = 0; t < 8; ++t) {
      uint8_t c1 = block[i1], c2 = block[i2];
      if (c1 != c2) return c1 > c2;
      uint16_t s1 = quadrant[i1], s2 = quadrant[i2];
      if (s1 != s2) return s1 > s2;
      i1++;
      i2++;
    }
    if (i1 >= nblock) i1 -= nblock;
    if (i2 >= nblock) i2 -= nblock;
    k -= 8;
    (*budget)--;
  } while (k >= 0);
  return false;
}

__device__ void main_simple_sort(uint32_t *ptr, const uint8_t *block, const uint16_t *quadrant, int nblock, int lo, int hi,
                                 int d, int *budget) {
  const int incs[14] = {1, 4, 13, 40, 121, 364, 1093, 3280, 9841, 29524, 88573, 265720, 797161, 2391484};
  int big_n = hi - lo + 1;
  if (big_n < 2) return;
  int hp = 0;
  while (incs[hp] < big_n) hp++;
  hp--;
  for (; hp >= 0; hp--) {
    int h = incs[hp];
    int i = lo + h;
    bool next_h = false;
    while (!next_h) {
      for (int rep = 0; rep < 3; ++rep) {
        if (i > hi) {
          next_h = true;
          break;
        }
        uint32_t v = ptr[i];
        int j = i;
        while (main_gtu(ptr[j - h] + (uint32_t)d, v + (uint32_t)d, block, quadrant, (uint32_t)nblock, budget)) {
          ptr[j] = ptr[j - h];
          j = j - h;
          if (j <= (lo + h - 1)) break;
        }
        ptr[j] = v;
        i++;
      }
      if (next_h) break;
      if (*budget < 0) return;
    }
  }
}

__device__ __forceinline__ uint8_t mmed3(uint8_t a, uint8_t b, uint8_t c) {
  uint8_t t;
  if (a > b) {
    t = a;
    a = b;
    b = t;
  }
  if (b > c) {
    b = c;
    if (a > b) b = a;
  }
  return b;
}

__device__ bool main_qsort3(uint32_t *ptr, const uint8_t *block, const uint16_t *quadrant, int nblock, int lo_st, int hi_st,
                            int d_st, int *budget) {
  int stack_lo[100], stack_hi[100], stack_d[100];
  int next_lo[3], next_hi[3], next_d[3];
  int sp = 0;
  stack_lo[sp] = lo_st;
  stack_hi[sp] = hi_st;
  stack_d[sp] = d_st;
  sp++;
  while (sp > 0) {
    if (sp >= 98) return false;
    sp--;
    int lo = stack_lo[sp], hi = stack_hi[sp], d = stack_d[sp];
    if (hi - lo < 20 || d > (N_RADIX + N_QSORT)) {
      main_simple_sort(ptr, block, quadrant, nblock, lo, hi, d, budget);
      if (*budget < 0) return true;
      continue;
    }
    int med = mmed3(block[ptr[lo] + d], block[ptr[hi] + d], block[ptr[(lo + hi) >> 1] + d]);
    int un_lo = lo, lt_lo = lo, un_hi = hi, gt_hi = hi;
    for (;;) {
      for (;;) {
        if (un_lo > un_hi) break;
        int n = (int)block[ptr[un_lo] + d] - med;
        if (n == 0) {
          swp(ptr[un_lo], ptr[lt_lo]);
          lt_lo++;
          un_lo++;
          continue;
        }
        if (n > 0) break;
        un_lo++;
      }
      for (;;) {
        if (un_lo > un_hi) break;
        int n = (int)block[ptr[un_hi] + d] - med;
        if (n == 0) {
          swp(ptr[un_hi], ptr[gt_hi]);
          gt_hi--;
          un_hi--;
          continue;
        }
        if (n < 0) break;
        un_hi--;
      }
      if (un_lo > un_hi) break;
      swp(ptr[un_lo], ptr[un_hi]);
      un_lo++;
      un_hi--;
    }
    if (un_hi != un_lo - 1) return false;
    if (gt_hi < lt_lo) {
      stack_lo[sp] = lo;
      stack_hi[sp] = hi;
      stack_d[sp] = d + 1;
      sp++;
      continue;
    }
    int n = (lt_lo - lo) < (un_lo - lt_lo) ? (lt_lo - lo) : (un_lo - lt_lo);
    for (int a = lo, b = un_lo - n, c = n; c > 0; a++, b++, c--) swp(ptr[a], ptr[b]);
    int m = (hi - gt_hi) < (gt_hi - un_hi) ? (hi - gt_hi) : (gt_hi - un_hi);
    for (int a = un_lo, b = hi - m + 1, c = m; c > 0; a++, b++, c--) swp(ptr[a], ptr[b]);
    n = lo + un_lo - lt_lo - 1;
    m = hi - (gt_hi - un_hi) + 1;
    next_lo[0] = lo;
    next_hi[0] = n;
    next_d[0] = d;
    next_lo[1] = m;
    next_hi[1] = hi;
    next_d[1] = d;
    next_lo[2] = n + 1;
    next_hi[2] = m - 1;
    next_d[2] = d + 1;
#define BZ_NSIZE(a) (next_hi[a] - next_lo[a])
#define BZ_NSWAP(a, b)      \
  {                         \
    int t_;                 \
    t_ = next_lo[a];        \
    next_lo[a] = next_lo[b]; \
    next_lo[b] = t_;        \
    t_ = next_hi[a];        \
    next_hi[a] = next_hi[b]; \
    next_hi[b] = t_;        \
    t_ = next_d[a];         \
    next_d[a] = next_d[b];  \
    next_d[b] = t_;         \
  }
    if (BZ_NSIZE(0) < BZ_NSIZE(1)) BZ_NSWAP(0, 1);
    if (BZ_NSIZE(1) < BZ_NSIZE(2)) BZ_NSWAP(1, 2);
    if (BZ_NSIZE(0) < BZ_NSIZE(1)) BZ_NSWAP(0, 1);
#undef BZ_NSIZE
#undef BZ_NSWAP
    for (int a = 0; a < 3; ++a) {
      stack_lo[sp] = next_lo[a];
      stack_hi[sp] = next_hi[a];
      stack_d[sp] = next_d[a];
      sp++;
    }
  }
  return true;
}

// _mainSort.  block has N_OVERSHOOT writable bytes after nblock; quadrant likewise.
__device__ bool main_sort(uint32_t *ptr, uint8_t *block, uint16_t *quadrant, uint32_t *ftab, int nblock, int *budget) {
  const uint32_t SETMASK = 1u << 21, CLEARMASK = ~(1u << 21);
  int running_order[256], copy_start[256], copy_end[256];
  bool big_done[256];
  for (int i = 65536; i >= 0; i--) ftab[i] = 0;
  int j = block[0] << 8;
  for (int i = nblock - 1; i >= 0; i--) {
    quadrant[i] = 0;
    j = (j >> 8) | ((int)block[i] << 8);
    ftab[j]++;
  }
  for (int i = 0; i < N_OVERSHOOT; i++) {
    block[nblock + i] = block[i];
    quadrant[nblock + i] = 0;
  }
  for (int i = 1; i <= 65536; i++) ftab[i] += ftab[i - 1];
  uint16_t s = (uint16_t)(block[0] << 8);
  for (int i = nblock - 1; i >= 0; i--) {
    s = (uint16_t)((s >> 8) | (block[i] << 8));
    j = (int)ftab[s] - 1;
    ftab[s] = (uint32_t)j;
    ptr[j] = (uint32_t)i;
  }
  for (int i = 0; i <= 255; i++) {
    big_done[i] = false;
    running_order[i] = i;
  }
#define BZ_BIGFREQ(b) (ftab[((b) + 1) << 8] - ftab[(b) << 8])
  {
    int h = 1;
    do h = 3 * h + 1;
    while (h <= 256);
    do {
      h = h / 3;
      for (int i = h; i <= 255; i++) {
        int vv = running_order[i];
        j = i;
        while (BZ_BIGFREQ(running_order[j - h]) > BZ_BIGFREQ(vv)) {
          running_order[j] = running_order[j - h];
          j = j - h;
          if (j <= (h - 1)) break;
        }
        running_order[j] = vv;
      }
    } while (h != 1);
  }
#undef BZ_BIGFREQ
  for (int i = 0; i <= 255; i++) {
    const int ss = running_order[i];
    for (j = 0; j <= 255; j++) {
      if (j != ss) {
        const int sb = (ss << 8) + j;
        if (!(ftab[sb] & SETMASK)) {
          int lo = (int)(ftab[sb] & CLEARMASK), hi = (int)(ftab[sb + 1] & CLEARMASK) - 1;
          if (hi > lo) {
            if (!main_qsort3(ptr, block, quadrant, nblock, lo, hi, N_RADIX, budget)) return false;
            if (*budget < 0) return true;
          }
        }
        ftab[sb] |= SETMASK;
      }
    }
    if (big_done[ss]) return false;
    for (j = 0; j <= 255; j++) {
      copy_start[j] = (int)(ftab[(j << 8) + ss] & CLEARMASK);
      copy_end[j] = (int)(ftab[(j << 8) + ss + 1] & CLEARMASK) - 1;
    }
    for (j = (int)(ftab[ss << 8] & CLEARMASK); j < copy_start[ss]; j++) {
      int k = (int)ptr[j] - 1;
      if (k < 0) k += nblock;
      uint8_t c1 = block[k];
      if (!big_done[c1]) ptr[copy_start[c1]++] = (uint32_t)k;
    }
    for (j = (int)(ftab[(ss + 1) << 8] & CLEARMASK) - 1; j > copy_end[ss]; j--) {
      int k = (int)ptr[j] - 1;
      if (k < 0) k += nblock;
      uint8_t c1 = block[k];
      if (!big_done[c1]) ptr[copy_end[c1]--] = (uint32_t)k;
    }
    if (!((copy_start[ss] - 1 == copy_end[ss]) || (copy_start[ss] == 0 && copy_end[ss] == nblock - 1))) return false;
    for (j = 0; j <= 255; j++) ftab[(j << 8) + ss] |= SETMASK;
    big_done[ss] = true;
    if (i < 255) {
      int bb_start = (int)(ftab[ss << 8] & CLEARMASK);
      int bb_size = (int)(ftab[(ss + 1) << 8] & CLEARMASK) - bb_start;
      int shifts = 0;
      if (bb_size > 0) {
        while ((bb_size >> shifts) > 65534) shifts++;
        for (j = bb_size - 1; j >= 0; j--) {
          int a2update = (int)ptr[bb_start + j];
          uint16_t q_val = (uint16_t)(j >> shifts);
          quadrant[a2update] = q_val;
          if (a2update < N_OVERSHOOT) quadrant[a2update + nblock] = q_val;
        }
      }
    }
  }
  return true;
}

// scratch per listed block: A (>= 7 MB): block copy [nblock + 34] | quadrant u16 [nblock + 34] | ftab u32 [65537];
// B (>= 3.7 MB): eclass u32 [nblock] | bhtab u32 [2 + nblock / 32 + 64]
__global__ void k_serial_sort(const uint8_t *__restrict__ blockbuf, const uint32_t *__restrict__ nblk,
                              const uint32_t *__restrict__ list, uint32_t n_list, uint32_t *__restrict__ SA,
                              uint32_t *__restrict__ origptr, uint8_t *__restrict__ scratchA, uint8_t *__restrict__ scratchB,
                              int *__restrict__ fail) {
  const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= n_list) return;
  const uint32_t bl = list[li];
  const int nblock = (int)nblk[bl];
  const uint8_t *src = blockbuf + (size_t)bl * BZ2E_BLKBYTES;
  uint32_t *ptr = SA + (size_t)bl * BZ2E_BSTRIDE;
  uint8_t *a = scratchA + (size_t)bl * BZ2E_BSTRIDE * 8;
  uint8_t *bsc = scratchB + (size_t)bl * BZ2E_BSTRIDE * 8;
  uint8_t *block = a;
  uint16_t *quadrant = (uint16_t *)(a + 1048576);
  uint32_t *ftab = (uint32_t *)(a + 1048576 + 2097152);
  uint32_t *eclass = (uint32_t *)bsc;
  uint32_t *bhtab = (uint32_t *)(bsc + 4194304);
  bool ok = true;
  if (nblock < 10000) {
    ok = fallback_sort(ptr, eclass, bhtab, src, nblock);
  } else {
    for (int i = 0; i < nblock; ++i) block[i] = src[i];
    int budget = nblock * ((30 - 1) / 3);  // workFactor 30 (bzip2_encoder.dart:45,909-915)
    ok = main_sort(ptr, block, quadrant, ftab, nblock, &budget);
    if (ok && budget < 0) ok = fallback_sort(ptr, eclass, bhtab, src, nblock);
  }
  int op = -1;
  if (ok)
    for (int i = 0; i < nblock; i++)
      if (ptr[i] == 0) {
        op = i;
        break;
      }
  if (op < 0) atomicExch(fail, 1);
  else origptr[bl] = (uint32_t)op;
}
}  // namespace serial

// h_cnt[i] != 0 marks the blocks of the batch that still have unresolved (= identical) rotations
static int serial_sort_blocks(const uint8_t *blockbuf, const uint32_t *nblk, const uint32_t *h_cnt, const uint32_t *h_n,
                              uint32_t nb, uint32_t *SA, uint32_t *origptr, void *scratchA, void *scratchB, void *scratchC,
                              cudaStream_t s) {
  (void)h_n;
  std::vector<uint32_t> h_list_v(nb + 1);
  uint32_t *h_list = h_list_v.data();
  uint32_t n_list = 0;
  for (uint32_t i = 0; i < nb; ++i)
    if (h_cnt[i]) h_list[n_list++] = i;
  h_list[nb] = 0;
  // scratchC (a u32 array of the sort, no longer needed): [0..nb) list, [nb] failure flag
  uint32_t *d_list = (uint32_t *)scratchC;
  int *d_fail = (int *)(d_list + nb);
  cudaMemcpyAsync(d_list, h_list, 4 * (size_t)n_list, cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(d_fail, 0, 4, s);
  B200Z_LAUNCH(serial::k_serial_sort, (n_list + 31) / 32, 32, 0, s, blockbuf, nblk, (const uint32_t *)d_list, n_list, SA,
               origptr, (uint8_t *)scratchA, (uint8_t *)scratchB, d_fail);
  int h_fail = 0;
  cudaMemcpyAsync(&h_fail, d_fail, 4, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess || h_fail) return -6;
  return 0;
}
