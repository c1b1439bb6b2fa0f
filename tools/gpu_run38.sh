mkdir -p gpurun_out
L=gpurun_out/r02L_kbench.log
: > $L
for lib in "" mb; do
  echo "lib '$lib'" >> $L
  if [ -n "$lib" ]; then export BSQ_LIB=$PWD/biseqt_b200/libbsqgpu_$lib.so; else unset BSQ_LIB; fi
  python tools/kbench.py c2_logodds 250000 4 >> $L 2>&1
  python tools/kbench.py c3_logodds 32768 3 >> $L 2>&1
  python tools/kbench.py c3_int 32768 3 >> $L 2>&1
  python tools/kbench.py c3_blocal_int 32768 3 >> $L 2>&1
  python tools/kbench.py c3_blocal_logodds 32768 3 >> $L 2>&1
  python tools/kbench.py c4_int 4096 3 >> $L 2>&1
done
grep -E "^lib|^c[234]|fill_ms" $L | cut -c1-120
