#!/bin/bash
# One single-GPU box session that produces everything profiles/ needs (run on the GPU box from the repo root):
#   tests -> per-stage times -> bench line -> ncu launch list of the bench command -> ncu --set full of one step -> GPU parity pin
# usage: bash tools/gpu_session.sh <tag>     (outputs under gpurun_out/<tag>_*)
T=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1
for m in 0x0 0x8 0xa 0xf; do timeout 300 python tools/stage_times.py 8 $m > $O/${T}_stage_$m.txt 2>&1; done
timeout 300 python tools/stage_times.py 8 0xa JzAzBz > $O/${T}_stage_jz.txt 2>&1
timeout 300 python tools/stage_times.py 8 0xa OKLAB > $O/${T}_stage_ok.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
timeout 600 python tests/golden/pin_expected_parity.py --gpu --out $O/expected_parity_gpu.json > $O/${T}_pin.log 2>&1
# ncu only after the same commands exited cleanly above
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${T}_launches_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-extra > $O/${T}_ncu_bench.log 2>&1
# the full report stays on the box unless it is small (gpurun_out/ is capped at 64 MiB): bring back the raw page and the
# SASS pages of the stencil kernels as CSV
timeout 900 ncu --set full --clock-control none --import-source on -f -o /tmp/${T}_full --launch-skip 48 --launch-count 24 \
    python tools/prof_step.py 16 0xa > $O/${T}_ncu_full.log 2>&1
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > $O/${T}_full_raw.csv 2>/dev/null
for k in k_prefilter k_canny_nms k_upsample2x_color_inverse k_color_forward_planar k_qt_blocks k_hysteresis; do
  ncu -i /tmp/${T}_full.ncu-rep --page source --csv --kernel-name regex:$k 2>/dev/null | gzip -9 > $O/${T}_sass_$k.csv.gz
done
sz=$(stat -c %s /tmp/${T}_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 40000000 ]; then cp /tmp/${T}_full.ncu-rep $O/; fi
du -sh $O
cat $O/${T}_pytest.log | tail -6
tail -n 2 $O/${T}_smoke.log
cat $O/${T}_stage_0xa.txt
grep -E "dct|idct|step" $O/${T}_stage_0x0.txt $O/${T}_stage_0xf.txt
cat $O/${T}_bench_n1.json | cut -c1-1500
for f in $O/${T}_bench_n1.err $O/${T}_pin.log $O/${T}_ncu_bench.log $O/${T}_ncu_full.log; do tail -n 2 $f; done
