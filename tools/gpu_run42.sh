mkdir -p gpurun_out
L=gpurun_out/r02P_kbench.log
: > $L
for tc in -1 3 4 6 8; do
  echo "target_chunks $tc" >> $L
  BSQ_TARGET_CHUNKS=$tc python tools/kbench.py c3_int 100000 3 >> $L 2>&1
  BSQ_TARGET_CHUNKS=$tc python tools/kbench.py c4_int 10000 3 >> $L 2>&1
done
grep -E "^target|^c[34]|total_ms" $L | cut -c1-120
