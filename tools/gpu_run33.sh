mkdir -p gpurun_out
L=gpurun_out/r02G_kbench.log
: > $L
python tools/kbench.py c5_int 1 3 >> $L 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "wide_band_batch or traceback_kernels or wide or full_size or forced or k3" > gpurun_out/r02G_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02G_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02G_launches_c5.csv python tools/kbench.py c5_int 1 1 > gpurun_out/r02G_ncu_c5.log 2>&1
tail -3 gpurun_out/r02G_pytest.log; cat $L; grep -i "walk" gpurun_out/r02G_launches_c5.csv | awk -F'","' '{print $5, $(NF)}'
