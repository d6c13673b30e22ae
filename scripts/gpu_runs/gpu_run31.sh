mkdir -p gpurun_out
BZE_MIB=64 BZE_CHECK=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_h_tables|k_m_mtf" -c 2 -o gpurun_out/r1_bz2enc_hot -f python scripts/bench_bz2enc.py > gpurun_out/bz2enc_ncu2.log 2>&1
ls -la gpurun_out/r1_bz2enc_hot.ncu-rep
