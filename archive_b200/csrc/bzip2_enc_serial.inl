// bzip2_enc_serial.inl -- placeholder, replaced below
static int serial_sort_blocks(const uint8_t *, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t, uint32_t *,
                              uint32_t *, void *, void *, void *, cudaStream_t) {
  return -6;
}
