for m in 1 2; do
 for cfg in "4096 512 1024 1" "512 4096 1024 1" "4096 1024 1024 1" "4096 256 1024 1" "4096 512 512 1" "4096 1024 1024 2" "1024 4096 1024 2"; do
  timeout 60 tools/bin/tc_gemm_bench 8192 256 256 $m $cfg 2>&1 | head -1
 done
done
