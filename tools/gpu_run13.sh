mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu"
$CMD > gpurun_out/r02m_plain.json 2> gpurun_out/r02m_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_c2.csv $CMD > gpurun_out/r02m_ncu_c2.log 2>&1
python tools/kbench.py c3_int 32768 2 > gpurun_out/r02m_plain_c3.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02m_launches_c3.csv python tools/kbench.py c3_int 32768 2 > gpurun_out/r02m_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsq_k1h_fill -s 3 -c 1 -o gpurun_out/r02m_k1h python tools/kbench.py c2_int 1000000 2 > gpurun_out/r02m_full_k1h.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bsq_k2_fill|bsq_k2_walk_coop" -s 2 -c 2 -o gpurun_out/r02m_k2 python tools/kbench.py c3_int 32768 2 > gpurun_out/r02m_full_k2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bsq_k2_fill" -s 1 -c 1 -o gpurun_out/r02m_k2_blocal python tools/kbench.py c3_blocal_int 32768 2 > gpurun_out/r02m_full_k2b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bsq_k2_fill" -s 1 -c 1 -o gpurun_out/r02m_k2_f64 python tools/kbench.py c3_logodds 32768 2 > gpurun_out/r02m_full_k2f.log 2>&1
BSQ_K3_CFG=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bsq_k3l_fill|bsq_k3_walk" -c 2 -o gpurun_out/r02m_k3l python tools/kbench.py c5_int 1 1 > gpurun_out/r02m_full_k3l.log 2>&1
ls -la gpurun_out/*.ncu-rep
