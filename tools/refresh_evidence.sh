#!/bin/bash
# One GPU call that regenerates the round's evidence into gpurun_out/<tag>_*: bench line, launch list, --set full captures of the
# forward / backward kernels (+ regions, traffic), DRAM bytes of the step in place, ops / configs tables, step split, phase probes.
# usage: tools/refresh_evidence.sh <tag>      (probe builds libmgw_probe_b.so / libmgw_probe_f.so are used if present)
tag=${1:-ev}
cd "$(dirname "$0")/.."
if [ -z "$ONLY_CAPTURE" ]; then
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err || tail -3 gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${tag}_launches.log 2>&1
fi
if [ -z "$SKIP_CAPTURE" ]; then
bash tools/ncu_capture.sh ${tag}c > /dev/null 2>&1            # one capture: the forward and the backward kernel of the third pass
python tools/ncu_regions.py gpurun_out/${tag}c_source.csv warp_fwd_pipe > gpurun_out/${tag}_regions_warp_fwd_pipe.txt 2>&1
python tools/ncu_regions.py gpurun_out/${tag}c_source.csv warp_bwd_tma > gpurun_out/${tag}_regions_warp_bwd_tile.txt 2>&1
python tools/ncu_traffic.py gpurun_out/${tag}c_raw.csv gpurun_out/${tag}c_raw.csv > gpurun_out/${tag}_traffic.json 2>&1
fi
[ -n "$ONLY_CAPTURE" ] && exit 0
bash tools/ncu_step_traffic.sh ${tag}_steptraffic > gpurun_out/${tag}_step_in_place.json 2>&1
timeout 300 python tools/bench_ops.py > gpurun_out/${tag}_ops.json 2> gpurun_out/${tag}_ops.err
timeout 600 python tools/bench_configs.py > gpurun_out/${tag}_configs.json 2> gpurun_out/${tag}_configs.err
timeout 300 python tools/step_split.py 32 > gpurun_out/${tag}_step_split.txt 2>&1
[ -f deep-online-video-stabilization_b200/libmgw_probe_b.so ] && MGW_SO_NAME=libmgw_probe_b.so python tools/probe_bwd.py > gpurun_out/${tag}_probe_bwd.txt 2>&1
[ -f deep-online-video-stabilization_b200/libmgw_probe_f.so ] && MGW_SO_NAME=libmgw_probe_f.so python tools/probe_fwd.py > gpurun_out/${tag}_probe_fwd.txt 2>&1
timeout 300 python tools/soak.py 300 5 > gpurun_out/${tag}_soak.txt 2>&1
ls -la gpurun_out/${tag}_* | head -40
