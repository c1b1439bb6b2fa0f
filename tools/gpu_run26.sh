mkdir -p gpurun_out /tmp/ncu
for lib in prev new; do
  if [ $lib = prev ]; then export BSQ_LIB=$PWD/biseqt_b200/libbsqgpu_prev.so; else unset BSQ_LIB; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsq_k3l_fill -c 1 -o /tmp/ncu/k3l_$lib python tools/kbench.py c5_int 1 1 > gpurun_out/r02z_full_$lib.log 2>&1
  python profiles/summarize_ncu.py /tmp/ncu/k3l_$lib.ncu-rep > gpurun_out/r02z_k3l_${lib}_ncu.txt 2>&1
  ncu -i /tmp/ncu/k3l_$lib.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02z_k3l_${lib}_source.csv.gz
done
head -8 gpurun_out/r02z_k3l_prev_ncu.txt; head -8 gpurun_out/r02z_k3l_new_ncu.txt
