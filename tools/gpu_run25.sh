mkdir -p gpurun_out
L=gpurun_out/r02y_kbench.log
: > $L
for lib in prev lin "" ; do
  echo "lib '$lib'" >> $L
  if [ -n "$lib" ]; then BSQ_LIB=$PWD/biseqt_b200/libbsqgpu_$lib.so python tools/kbench.py c5_int 1 3 >> $L 2>&1; else python tools/kbench.py c5_int 1 3 >> $L 2>&1; fi
done
BSQ_K3_CFG=0 python tools/k3prof.py c5_int > gpurun_out/r02y_k3prof.txt 2>&1
for n in 16 64 256 1024; do
  echo "c1n $n as bands" >> $L; BSQ_STD_BAND_MAX=1000000 python tools/kbench.py c1n $n 3 >> $L 2>&1
  echo "c1n $n on K1" >> $L; BSQ_NO_STD_BAND=1 python tools/kbench.py c1n $n 3 >> $L 2>&1
done
cat $L; tail -4 gpurun_out/r02y_k3prof.txt
