mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not c5 and not big_ragged and not ragged_reads" > gpurun_out/r02h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_pytest.log
L=gpurun_out/r02h_kbench.log
: > $L
for k in c3_int c3_logodds c3_blocal_int c3_blocal_logodds; do python tools/kbench.py $k 32768 4 >> $L 2>&1; done
python tools/kbench.py c3_int 32768 3 notrace >> $L 2>&1
python tools/kbench.py c4_int 4096 3 >> $L 2>&1
python tools/kbench.py c4_int 4096 3 notrace >> $L 2>&1
BSQ_K2_CFG=2 python tools/kbench.py c3_int 32768 3 >> $L 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02h_launches_c3.csv python tools/kbench.py c3_int 32768 2 > gpurun_out/r02h_ncu_c3.log 2>&1
tail -4 gpurun_out/r02h_pytest.log; cat $L; grep -E "walk|k2_fill|pack_tr" gpurun_out/r02h_launches_c3.csv | awk -F'","' '{print $5, $NF}' | head
