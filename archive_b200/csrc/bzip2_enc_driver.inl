// bzip2_enc_driver.inl -- host orchestration of the device BZip2 encoder.  Included inside namespace b200z::bz2e.

namespace {
struct Carver {
  uint8_t *base;
  size_t off = 0;
  explicit Carver(void *p) : base((uint8_t *)p) {}
  template <class T>
  T *take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T *r = base ? (T *)(base + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
};
struct Bufs {
  // whole input
  uint32_t *t_head, *t_tail, *pre, *t_sum, *sub_pre, *n_blocks;
  uint16_t *sub_sum;
  unsigned long long *G, *out_bytes;
  BlkInfo *blk;
  EncState *st;
  // per batch
  uint8_t *blockbuf, *lsym, *mtfpos, *selector, *selmtf, *lens;
  uint32_t *inuse, *nblk, *cnt, *cnt_next, *origptr, *block_crc, *part_crc, *part_len;
  unsigned long long *keysA, *keysB, *totals, *bit_off;
  uint32_t *valsA, *valsB, *slotsA, *slotsB, *SA, *rank, *tile_hist, *tile_active, *act_off, *tile_cnt, *tile_off, *nmtf,
      *mtf_freq, *codes, *tile_bitoff;
  int *tile_lastflag, *carry, *lastocc, *tile_nzlast;
  uint16_t *mtfv;
  HInfo *hinfo;
};
size_t carve(void *ws, uint32_t nt, uint32_t max_blocks, uint32_t nb, Bufs &b) {
  Carver c(ws);
  b.t_head = c.take<uint32_t>(nt);
  b.t_tail = c.take<uint32_t>(nt);
  b.pre = c.take<uint32_t>(nt);
  b.t_sum = c.take<uint32_t>(nt);
  b.sub_sum = c.take<uint16_t>((size_t)nt * SUBS);
  b.sub_pre = c.take<uint32_t>((size_t)nt * SUBS + 1);
  b.G = c.take<unsigned long long>((size_t)nt + 1);
  b.blk = c.take<BlkInfo>(max_blocks);
  b.n_blocks = c.take<uint32_t>(2);
  b.st = c.take<EncState>(1);
  b.out_bytes = c.take<unsigned long long>(1);
  const size_t e = (size_t)nb * BZ2E_BSTRIDE;
  b.blockbuf = c.take<uint8_t>((size_t)nb * BZ2E_BLKBYTES + 16);
  b.inuse = c.take<uint32_t>((size_t)nb * 8);
  b.nblk = c.take<uint32_t>(nb);
  b.cnt = c.take<uint32_t>(nb);
  b.cnt_next = c.take<uint32_t>(nb);
  b.origptr = c.take<uint32_t>(nb);
  b.block_crc = c.take<uint32_t>(nb);
  b.part_crc = c.take<uint32_t>((size_t)nb * CRC_PARTS);
  b.part_len = c.take<uint32_t>((size_t)nb * CRC_PARTS);
  b.keysA = c.take<unsigned long long>(e);
  b.keysB = c.take<unsigned long long>(e);
  b.valsA = c.take<uint32_t>(e);
  b.valsB = c.take<uint32_t>(e);
  b.slotsA = c.take<uint32_t>(e);
  b.slotsB = c.take<uint32_t>(e);
  b.SA = c.take<uint32_t>(e);
  b.rank = c.take<uint32_t>(e);
  b.tile_hist = c.take<uint32_t>((size_t)nb * 256 * NT);
  b.tile_lastflag = c.take<int>((size_t)nb * NT);
  b.tile_active = c.take<uint32_t>((size_t)nb * NT);
  b.carry = c.take<int>((size_t)nb * NT);
  b.act_off = c.take<uint32_t>((size_t)nb * NT);
  b.totals = c.take<unsigned long long>(2);
  b.lsym = c.take<uint8_t>(e);
  b.mtfpos = c.take<uint8_t>(e + 16);
  b.lastocc = c.take<int>((size_t)nb * NT * 256);
  b.tile_nzlast = c.take<int>((size_t)nb * NT);
  b.tile_cnt = c.take<uint32_t>((size_t)nb * NT);
  b.tile_off = c.take<uint32_t>((size_t)nb * NT);
  b.nmtf = c.take<uint32_t>(nb);
  b.mtf_freq = c.take<uint32_t>((size_t)nb * BZ_MAX_ALPHA);
  b.mtfv = c.take<uint16_t>(e);
  b.selector = c.take<uint8_t>((size_t)nb * SEL_STRIDE);
  b.selmtf = c.take<uint8_t>((size_t)nb * SEL_STRIDE);
  b.lens = c.take<uint8_t>((size_t)nb * BZ_N_GROUPS * BZ_MAX_ALPHA);
  b.codes = c.take<uint32_t>((size_t)nb * BZ_N_GROUPS * BZ_MAX_ALPHA);
  b.hinfo = c.take<HInfo>(nb);
  b.tile_bitoff = c.take<uint32_t>((size_t)nb * (NET + 1));
  b.bit_off = c.take<unsigned long long>(nb);
  return (c.off + 255) & ~(size_t)255;
}
}  // namespace

size_t bound(size_t n) { return n + n / 32 + 8192; }

Plan plan(size_t n, size_t mem_budget) {
  Plan p;
  p.n_tiles = (uint32_t)((n + TI - 1) / TI);
  // a block consumes at least (899 981 - 4) / 5 * 4 input bytes (every run of 4 becomes 5 bytes)
  p.max_blocks = (uint32_t)(n / 700000 + 2);
  uint32_t want = (uint32_t)(n / 890000 + 2);
  if (want > 256) want = 256;
  Bufs b;
  for (;;) {
    p.batch = want;
    p.ws_bytes = carve(nullptr, p.n_tiles ? p.n_tiles : 1, p.max_blocks, want, b);
    if (p.ws_bytes <= mem_budget || want == 1) break;
    want = want > 8 ? want / 2 : want - 1;
  }
  return p;
}

#define BZ2E_CK(x)                         \
  do {                                     \
    cudaError_t e_ = (x);                  \
    if (e_ != cudaSuccess) return -6;      \
  } while (0)

int encode_device(const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap, void *ws, const Plan &p, size_t *out_len,
                  Stats *stats, void *stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  Bufs b;
  const uint32_t nt = p.n_tiles ? p.n_tiles : 1;
  carve(ws, nt, p.max_blocks, p.batch, b);
  uint32_t *out32 = (uint32_t *)d_out;
  if (out_cap < 14 + 8 || ((uintptr_t)d_out & 3)) return -3;
  const size_t zero_bytes = out_cap & ~(size_t)3;
  BZ2E_CK(cudaMemsetAsync(d_out, 0, zero_bytes, s));
  EncState st0{32, 0, 0};
  BZ2E_CK(cudaMemcpyAsync(b.st, &st0, sizeof st0, cudaMemcpyHostToDevice, s));
  Stats stt{0, 0, 0, 0};
  uint32_t h_nb[2] = {0, 0};
  std::vector<BlkInfo> h_blk_v;
  BlkInfo *h_blk = nullptr;
  if (n > 0) {
    const uint32_t n32 = (uint32_t)n;
    B200Z_LAUNCH(k_e_tile_info, nt, 256, 0, s, d_in, n32, b.t_head, b.t_tail);
    B200Z_LAUNCH(k_e_tile_pre, 1, 1024, 0, s, d_in, nt, b.t_head, b.t_tail, b.pre);
    B200Z_LAUNCH(k_e_tile_emit<false>, nt, 256, 0, s, d_in, n32, 0u, b.pre, b.t_sum, b.sub_sum, b.sub_pre,
                 (const unsigned long long *)nullptr, (const BlkInfo *)nullptr, 0u, 0u, (uint8_t *)nullptr,
                 (uint32_t *)nullptr);
    B200Z_LAUNCH(k_scan_u32_u64, 1, 1024, 0, s, b.t_sum, nt, b.G);
    B200Z_LAUNCH(k_e_cut, 1, 32, 0, s, d_in, n32, nt, b.sub_sum, b.sub_pre, b.G, b.blk, p.max_blocks, b.n_blocks);
    BZ2E_CK(cudaMemcpyAsync(h_nb, b.n_blocks, 8, cudaMemcpyDeviceToHost, s));
    BZ2E_CK(cudaStreamSynchronize(s));
    if (h_nb[1] || h_nb[0] == 0) return -6;
    h_blk_v.resize(h_nb[0]);
    h_blk = h_blk_v.data();
    BZ2E_CK(cudaMemcpyAsync(h_blk, b.blk, sizeof(BlkInfo) * h_nb[0], cudaMemcpyDeviceToHost, s));
    BZ2E_CK(cudaStreamSynchronize(s));
  }
  const uint32_t nblocks = h_nb[0];
  stt.n_blocks = nblocks;
  int rc = 0;
  std::vector<uint32_t> h_n_v(p.batch), h_cnt_v(p.batch);
  uint32_t *h_n = h_n_v.data(), *h_cnt = h_cnt_v.data();
  for (uint32_t lo = 0; lo < nblocks && rc == 0; lo += p.batch) {
    const uint32_t nb = (nblocks - lo < p.batch) ? nblocks - lo : p.batch;
    const uint32_t hi = lo + nb;
    uint32_t max_n = 0;
    for (uint32_t i = 0; i < nb; ++i) {
      h_n[i] = h_blk[lo + i].nblock;
      if (h_n[i] > max_n) max_n = h_n[i];
    }
    cudaMemcpyAsync(b.nblk, h_n, 4 * nb, cudaMemcpyHostToDevice, s);
    cudaMemsetAsync(b.inuse, 0, (size_t)nb * 32, s);
    cudaMemsetAsync(b.mtf_freq, 0, (size_t)nb * BZ_MAX_ALPHA * 4, s);
    // front end: RLE1 bytes + CRC of the batch's blocks
    {
      const uint32_t t0 = h_blk[lo].start / TI, t1 = (h_blk[hi - 1].end + TI - 1) / TI;
      B200Z_LAUNCH(k_e_fill_head, dim3(8, nb), 256, 0, s, d_in, (uint32_t)n, b.blk, lo, b.blockbuf, b.inuse);
      B200Z_LAUNCH(k_e_tile_emit<true>, t1 - t0, 256, 0, s, d_in, (uint32_t)n, t0, b.pre, (uint32_t *)nullptr,
                   (uint16_t *)nullptr, (uint32_t *)nullptr, b.G, b.blk, lo, hi, b.blockbuf, b.inuse);
      B200Z_LAUNCH(k_e_crc_part, dim3(CRC_PARTS, nb), 256, 0, s, d_in, b.blk, lo, b.part_crc, b.part_len);
      B200Z_LAUNCH(k_e_crc_final, (nb + 63) / 64, 64, 0, s, b.part_crc, b.part_len, nb, b.block_crc);
    }
    // rotation order
    uint32_t max_cnt = max_n;
    {
      const uint32_t ntl = (max_cnt + TS - 1) / TS;
      B200Z_LAUNCH(k_s_init, dim3(ntl, nb), 256, 0, s, b.blockbuf, b.nblk, b.keysA, b.valsA, b.slotsA, b.cnt);
    }
    unsigned long long *kA = b.keysA, *kB = b.keysB;
    uint32_t *vA = b.valsA, *vB = b.valsB, *sA = b.slotsA, *sB = b.slotsB, *cA = b.cnt, *cB = b.cnt_next;
    uint32_t h = 5;
    bool ties = false;
    for (;;) {
      const uint32_t ntl = (max_cnt + TS - 1) / TS;
      for (uint32_t pass = 0; pass < 5; ++pass) {
        B200Z_LAUNCH(k_s_hist, dim3(ntl, nb), 256, 0, s, kA, cA, pass * 8, b.tile_hist);
        B200Z_LAUNCH(k_s_scan, nb, 256, 0, s, b.tile_hist, cA);
        B200Z_LAUNCH(k_s_scatter, dim3(ntl, nb), 256, 0, s, kA, vA, kB, vB, cA, pass * 8, b.tile_hist);
        unsigned long long *tk = kA;
        kA = kB;
        kB = tk;
        uint32_t *tv = vA;
        vA = vB;
        vB = tv;
      }
      cudaMemsetAsync(b.totals, 0, 16, s);
      B200Z_LAUNCH(k_s_groups_count, dim3(ntl, nb), 256, 0, s, kA, cA, b.tile_lastflag, b.tile_active);
      B200Z_LAUNCH(k_s_block_scan, nb, 512, 0, s, b.tile_lastflag, b.tile_active, cA, cB, b.carry, b.act_off, b.totals);
      B200Z_LAUNCH(k_s_update, dim3(ntl, nb), 256, 0, s, kA, vA, sA, cA, b.carry, b.SA, b.rank, b.origptr);
      unsigned long long h_tot[2];
      BZ2E_CK(cudaMemcpyAsync(h_tot, b.totals, 16, cudaMemcpyDeviceToHost, s));
      BZ2E_CK(cudaStreamSynchronize(s));
      stt.rounds++;
#ifdef B200Z_EMU
      if (getenv("BZ2E_DEBUG")) fprintf(stderr, "round %u h=%u active=%llu max=%llu\n", stt.rounds, h, h_tot[0], h_tot[1]);
#endif
      if (h_tot[0] == 0) break;
      if (h >= max_n) {
        ties = true;
        break;
      }
      B200Z_LAUNCH(k_s_build, dim3(ntl, nb), 256, 0, s, kA, vA, sA, cA, b.act_off, b.rank, b.nblk, h, kB, vB, sB);
      {
        unsigned long long *tk = kA;
        kA = kB;
        kB = tk;
        uint32_t *tv = vA;
        vA = vB;
        vB = tv;
        uint32_t *ts = sA;
        sA = sB;
        sB = ts;
        uint32_t *tc = cA;
        cA = cB;
        cB = tc;
      }
      max_cnt = (uint32_t)h_tot[1];
      h *= 2;
    }
    if (ties) {
      // blocks whose rotations are not all distinct: the reference's order among equal rotations is an artefact of
      // its sort, so those blocks run the serial restatement
      BZ2E_CK(cudaMemcpyAsync(h_cnt, cB, 4 * nb, cudaMemcpyDeviceToHost, s));
      BZ2E_CK(cudaStreamSynchronize(s));
      int r2 = serial_sort_blocks(b.blockbuf, b.nblk, h_cnt, h_n, nb, b.SA, b.origptr, (void *)kA, (void *)kB, (void *)vA,
                                  s);
      if (r2 != 0) {
        rc = r2;
        break;
      }
      for (uint32_t i = 0; i < nb; ++i)
        if (h_cnt[i]) stt.n_serial_blocks++;
    }
    // MTF + RUNA/RUNB
    {
      const uint32_t ntl = (max_n + TS - 1) / TS;
      B200Z_LAUNCH(k_m_lsym, dim3(ntl, nb), 256, 0, s, b.blockbuf, b.SA, b.nblk, b.inuse, b.lsym, b.lastocc);
      B200Z_LAUNCH(k_m_scan_last, nb, 256, 0, s, b.lastocc, b.nblk);
      B200Z_LAUNCH(k_m_mtf, dim3((ntl + MTF_CPB - 1) / MTF_CPB, nb), MTF_CPB, 0, s, b.lsym, b.lastocc, b.nblk, b.inuse, b.mtfpos, b.tile_nzlast);
      B200Z_LAUNCH(k_m_zr<false>, dim3(ntl, nb), 256, 0, s, b.mtfpos, b.tile_nzlast, b.nblk, b.tile_cnt,
                   (const uint32_t *)nullptr, (const uint32_t *)nullptr, (const uint32_t *)nullptr, (uint16_t *)nullptr,
                   (uint32_t *)nullptr);
      B200Z_LAUNCH(k_m_zr_scan, nb, 512, 0, s, b.tile_cnt, b.nblk, b.tile_off, b.nmtf);
      B200Z_LAUNCH(k_m_zr<true>, dim3(ntl, nb), 256, 0, s, b.mtfpos, b.tile_nzlast, b.nblk, (uint32_t *)nullptr, b.tile_off,
                   b.nmtf, b.inuse, b.mtfv, b.mtf_freq);
    }
    // coding tables, offsets, emission
    B200Z_LAUNCH(k_h_tables, nb, 512, 0, s, b.mtfv, b.nmtf, b.mtf_freq, b.inuse, b.selector, b.selmtf, b.lens, b.codes,
                 b.hinfo, b.tile_bitoff);
    B200Z_LAUNCH(k_h_offsets, 1, 32, 0, s, b.hinfo, b.block_crc, nb, 0u, b.st, b.bit_off, out32);
    EncState h_st;
    BZ2E_CK(cudaMemcpyAsync(&h_st, b.st, sizeof h_st, cudaMemcpyDeviceToHost, s));
    BZ2E_CK(cudaStreamSynchronize(s));
    if ((h_st.bitpos + 80 + 7) / 8 + 8 > out_cap) {
      rc = -3;
      break;
    }
    B200Z_LAUNCH(k_h_emit_header, nb, 256, 0, s, b.hinfo, b.bit_off, b.block_crc, b.origptr, b.inuse, b.selmtf, b.lens,
                 out32);
    B200Z_LAUNCH(k_h_emit_data, dim3(NET, nb), 256, 0, s, b.mtfv, b.nmtf, b.hinfo, b.bit_off, b.tile_bitoff, b.selector,
                 b.lens, b.codes, out32);
  }
  if (rc == -3) {
    *out_len = bound(n);
    return -3;
  }
  if (rc != 0) return rc;
  B200Z_LAUNCH(k_h_finish, 1, 32, 0, s, b.st, out32, b.out_bytes);
  unsigned long long h_bytes = 0;
  BZ2E_CK(cudaMemcpyAsync(&h_bytes, b.out_bytes, 8, cudaMemcpyDeviceToHost, s));
  BZ2E_CK(cudaStreamSynchronize(s));
  BZ2E_CK(cudaGetLastError());
  *out_len = (size_t)h_bytes;
  if (stats) *stats = stt;
  return 0;
}
