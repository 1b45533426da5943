for d in ${DBGS:-0 7 3 5 6}; do echo "DBG=$d"; DEVIT_GEMM_DBG=$d timeout 60 python tools/sweep_gemm.py dense ${SHAPES:-qkv} | sed 's/|.*192\/1/ 192\/1/' ; done
