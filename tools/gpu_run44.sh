mkdir -p gpurun_out
L=gpurun_out/r02R_kbench.log
: > $L
for lib in "" w3 w4; do
  echo "lib '$lib'" >> $L
  if [ -n "$lib" ]; then export BSQ_LIB=$PWD/biseqt_b200/libbsqgpu_$lib.so; else unset BSQ_LIB; fi
  python tools/kbench.py c3_blocal_logodds 32768 3 >> $L 2>&1
  python tools/kbench.py c3_blocal_logodds 100000 3 >> $L 2>&1
done
grep -E "^lib|^c3|_ms" $L | cut -c1-120
