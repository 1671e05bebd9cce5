# One GPU call that regenerates the round-2 set under profiles/ for the current build (see profiles/README.md).
set -x
mkdir -p gpurun_out
# 1. every launch of two steady-state steps of configs[1] (mixed), serialised (graphs, side stream, PDL off)
DISPNET_B200_GRAPHS=0 DISPNET_B200_SIDE_STREAM=0 DISPNET_B200_PHASE_STREAMS=0 DN_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 2 --warmup 3 --minimal > gpurun_out/r2_launches.log 2>&1
tail -1 gpurun_out/r2_launches.log
DISPNET_B200_GRAPHS=0 DISPNET_B200_SIDE_STREAM=0 DISPNET_B200_PHASE_STREAMS=0 DN_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_tc32_ncu.csv python bench.py --precision tc32 --steps 2 --warmup 3 --minimal > gpurun_out/r2_launches_tc32.log 2>&1
# 2. full captures of one launch each
# (the .ncu-rep files are excerpted on the box and deleted: gpurun copies at most 64 MiB back)
cap() { timeout 300 ncu --set full --import-source on --clock-control none -k regex:$1 -s $2 -c 1 -f -o gpurun_out/r2_full_$3 ${@:4} > gpurun_out/r2_full_$3.log 2>&1; python tools/ncu_excerpt.py gpurun_out/r2_full_$3.ncu-rep > gpurun_out/r2_full_$3.txt 2>/dev/null; rm -f gpurun_out/r2_full_$3.ncu-rep; }
cap igemm_tc 2 igemm_tc_feat27 python tools/prof_conv.py feat27
cap igemm_tc 2 igemm_tc_feat10 python tools/prof_conv.py feat10
cap igemm_halo 2 igemm_halo_iconv0 python tools/prof_conv.py iconv0
cap igemm_halo 2 igemm_halo_feat0 python tools/prof_conv.py feat0
cap wgrad_tc 2 wgrad_tc_iconv0 python tools/prof_conv.py iconv0
cap igemm_tc 2 igemm_tc_upconv0 python tools/prof_conv.py upconv0
cap igemm_halo 2 igemm_halo_feat3 python tools/prof_conv.py feat3
cap wgrad_tc 2 wgrad_tc_feat27 python tools/prof_conv.py feat27
cap wgrad_tc 2 wgrad_tc_feat3 python tools/prof_conv.py feat3
DISPNET_B200_PRECISION=tc32 cap igemm_tc 2 igemm_tc32_feat27 python tools/prof_conv.py feat27
DISPNET_B200_PRECISION=tc32 cap split_bf16 2 split_bf16_feat3 python tools/prof_conv.py feat3
cap photo_batch_fwd 1 photo_batch_fwd python tools/prof_loss.py 1
cap photo_batch_bwd 1 photo_batch_bwd python tools/prof_loss.py 1
cap area_pyramid 1 area_pyramid python tools/prof_loss.py 1
cap smooth_fwd 4 smooth_fwd python tools/prof_loss.py 1
cap smooth_bwd 4 smooth_bwd python tools/prof_loss.py 1
cap dl_partial 1 dl_partial python tools/prof_loss.py 1
cap dl_bwd 1 dl_bwd python tools/prof_loss.py 1
ls -la gpurun_out/r2_full_*.txt
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_loss_launches.csv python tools/prof_loss.py 1 > /dev/null 2>&1
# 3. numbers (never under a profiler)
timeout 200 python tools/prof_loss.py 20 > gpurun_out/r2_loss_timing.txt 2>&1; cat gpurun_out/r2_loss_timing.txt
timeout 900 python bench.py --dump gpurun_out/r2_layers.txt > gpurun_out/r2_bench_final.log 2> gpurun_out/r2_bench_final.err; tail -c 600 gpurun_out/r2_bench_final.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1; tail -c 400 gpurun_out/r2_bench_ref.log
timeout 600 python bench.py --precision tc32 --no-extras --no-cpu --steps 20 --warmup 5 --dump gpurun_out/r2_layers_tc32.txt > gpurun_out/r2_bench_tc32.log 2>&1; tail -c 300 gpurun_out/r2_bench_tc32.log
timeout 600 python tools/bench_configs.py 32 0,2,4 2>&1 | grep "configs\[" > gpurun_out/r2_configs.txt; timeout 300 python tools/bench_configs.py 16 3 2>&1 | grep "configs\[" >> gpurun_out/r2_configs.txt; cat gpurun_out/r2_configs.txt
timeout 200 python tools/grad_parity.py 4 tc32 mixed > gpurun_out/r2_grad_parity.txt 2>&1; tail -2 gpurun_out/r2_grad_parity.txt
