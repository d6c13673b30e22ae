mkdir -p gpurun_out
for cfg in "8 0" "16 0" "4 0"; do set -- $cfg; echo "== upw=$1"; B200Z_UPW=$1 timeout 600 python scripts/dbg_pieces.py 2>&1 | tail -9; done
