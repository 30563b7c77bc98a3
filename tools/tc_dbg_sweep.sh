# timing experiments on the tensor-core kernel (results are numerically meaningless with SACX_TC_DBG != 0)
for d in 0 4 12 20 36 68 124; do
  SACX_TC_DBG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sacx_tc" -s 30 -c 15 --csv --log-file gpurun_out/dbg_$d.csv python bench.py --workload dp --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
  echo "dbg=$d"; python tools/launch_summary.py gpurun_out/dbg_$d.csv 15 -v | grep sacx_tc | head -16 | awk '{printf "%s ", $NF} END {print ""}'
done
