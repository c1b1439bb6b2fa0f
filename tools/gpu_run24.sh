mkdir -p gpurun_out
L=gpurun_out/r02x_kbench.log
: > $L
for i in 1 2; do
  echo "prev lib" >> $L; BSQ_LIB=$PWD/biseqt_b200/libbsqgpu_prev.so python tools/kbench.py c5_int 1 3 >> $L 2>&1
  echo "new lib" >> $L; python tools/kbench.py c5_int 1 3 >> $L 2>&1
done
for n in 16 64 256 1024 4096; do
  echo "c1n $n as bands" >> $L; BSQ_STD_BAND_MAX=1000000 python tools/kbench.py c1n $n 3 >> $L 2>&1
  echo "c1n $n on K1" >> $L; BSQ_NO_STD_BAND=1 python tools/kbench.py c1n $n 3 >> $L 2>&1
done
cat $L
