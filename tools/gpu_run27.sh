mkdir -p gpurun_out
L=gpurun_out/r02A_kbench.log
: > $L
echo "prev lib" >> $L; BSQ_LIB=$PWD/biseqt_b200/libbsqgpu_prev.so python tools/kbench.py c5_int 1 3 >> $L 2>&1
python tools/kbench.py c5_int 1 3 >> $L 2>&1
python tools/kbench.py c5_logodds 1 2 >> $L 2>&1
python tools/kbench.py c5_int 10 2 >> $L 2>&1
python tools/kbench.py c5s_int 64 2 >> $L 2>&1
python tools/kbench.py c5s_logodds 64 2 >> $L 2>&1
for n in 2048 4096 8192; do
  echo "c1n $n as bands" >> $L; BSQ_STD_BAND_MAX=1000000 python tools/kbench.py c1n $n 3 >> $L 2>&1
  echo "c1n $n on K1" >> $L; BSQ_NO_STD_BAND=1 python tools/kbench.py c1n $n 3 >> $L 2>&1
done
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "traceback_kernels or bench_shapes or wide or full_size or std_as_band or forced or k3" > gpurun_out/r02A_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02A_pytest.log
tail -3 gpurun_out/r02A_pytest.log; cat $L
