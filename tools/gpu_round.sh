#!/bin/bash
# one gpurun call that produces everything profiles/ holds for a round: parity tests, smoke, both bench arms, the batched
# workloads, the ncu launch list of the bench command and one ncu --set full capture of the six level kernels
TAG=${1:-r02d}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $O/smi.txt 2>&1; nproc >> $O/smi.txt
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 40 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err
python bench.py --steps 10 --warmup 3 --workload c5 --no-cpu > $O/bench_c5_n1.json 2> $O/bench_c5.err
python bench.py --steps 10 --warmup 3 --workload c2b8 --no-cpu > $O/bench_c2b8.json 2> $O/bench_c2b8.err
python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err
python tools/prof_seq.py > $O/sequence.txt 2>&1
python tools/prof_swt.py > $O/swt_levels.txt 2>&1
python tools/time_inv_variants.py > $O/inv_variants.txt 2>&1
for cfg in "2 4096" "8 2048" "8 4096"; do set -- $cfg; B=$1 N=$2 REPS=4 python tools/tma_inv_stress.py 2>&1 | tail -1; done > $O/tma_inv_stress.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 12 -c 6 -o $O/ncu_c2 python tools/prof_fwdinv.py 3 > $O/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 6 -o $O/ncu_b16 python tools/prof_batch.py > $O/ncu_b16.log 2>&1
# the other kernel families: reductions / thresholds, SWT and non-separable level kernels
ncu --set full --clock-control none --import-source on -k regex:"k_reduce|k_threshold" -c 6 -o $O/ncu_elem python tools/prof_elem.py > $O/ncu_elem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_swt|k_nonsep" -s 8 -c 16 -o $O/ncu_c3c4 python tools/prof_c3c4.py > $O/ncu_c3c4.log 2>&1
for r in c2 b16 elem c3c4; do ncu -i $O/ncu_$r.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > $O/ncu_${r}_summary.txt; done
# per-launch DRAM bytes of the level-1 kernels at HEAD (bench.py -> roofline.traffic), then drop the big reports:
# gpurun only brings back 64 MiB
ncu -i $O/ncu_c2.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_traffic.py > $O/traffic.json
for k in 0 5; do
  ncu -i $O/ncu_b16.ncu-rep --page source --csv --print-source sass --launch-skip $k --launch-count 1 > $O/src_$k.csv 2>/dev/null
  python tools/ncu_source_stalls.py $O/src_$k.csv 30 > $O/stalls_b16_$k.txt; rm -f $O/src_$k.csv
done
cuobjdump -sass -fun '_ZN4pdwt14k_fwd2d_streamILi14ELb1EEEvNS_9FwdParamsIXT_EEE' pdwt_b200/_build/pdwt_stream.o | grep -E "UTMALDG|SYNCS|LDGSTS|FFMA2" | sed 's#/\* 0x[0-9a-f]* \*/##' | awk '{c[$2]++} END{for(k in c) print k, c[k]}' > $O/sass_fwd_opcodes.txt
cuobjdump -sass -fun '_ZN4pdwt11k_inv2d_tmaILi14EEEvNS_12InvTmaParamsIXT_EEE' pdwt_b200/_build/pdwt_stream.o | grep -E "UTMALDG|SYNCS|LDGSTS|FFMA2|LDS" | sed 's#/\* 0x[0-9a-f]* \*/##' | awk '{c[$2]++} END{for(k in c) print k, c[k]}' > $O/sass_inv_tma_opcodes.txt
rm -f $O/*.ncu-rep
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cut -c1-300 $O/bench_ref.json; cut -c1-300 $O/bench_ours.json; cat $O/sequence.txt | head -5
