mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu"
$CMD > gpurun_out/r02m_plain.json 2> gpurun_out/r02m_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_c2.csv $CMD > gpurun_out/r02m_ncu_c2.log 2>&1
python tools/kbench.py c3_int 32768 2 > gpurun_out/r02m_plain_c3.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02m_launches_c3.csv python tools/kbench.py c3_int 32768 2 > gpurun_out/r02m_ncu_c3.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; sk=$3; cn=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $sk -c $cn -o /tmp/ncu/$name "$@" > gpurun_out/r02m_full_$name.log 2>&1
  python profiles/summarize_ncu.py /tmp/ncu/$name.ncu-rep > gpurun_out/r02m_${name}_ncu.txt 2>&1
}
cap k1h bsq_k1h_fill 3 1 python tools/kbench.py c2_int 1000000 2
cap k2 "bsq_k2_fill|bsq_k2_walk_coop" 2 2 python tools/kbench.py c3_int 32768 2
cap k2_blocal bsq_k2_fill 1 1 python tools/kbench.py c3_blocal_int 32768 2
cap k2_f64 bsq_k2_fill 1 1 python tools/kbench.py c3_logodds 32768 2
BSQ_K3_CFG=0 cap k3l "bsq_k3l_fill|bsq_k3_walk" 0 2 python tools/kbench.py c5_int 1 1
ncu -i /tmp/ncu/k3l.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02m_k3l_source.csv.gz
ncu -i /tmp/ncu/k2.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02m_k2_source.csv.gz
cp /tmp/ncu/k1h.ncu-rep gpurun_out/r02m_k1h.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | head -40
