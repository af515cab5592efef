# Phase timeline of the tcgen05 GEMM CTAs (tools/tc_timeline.py); build the instrumented library HERE first:
#   python tools/tc_timeline.py --build && gpurun -- 'bash tools/gpu_timeline.sh'
mkdir -p gpurun_out
export SLIDE_B200_LIB=$PWD/slide_b200/libslide_b200_tl.so
timeout 300 python tools/tc_timeline.py lat 256 ${LAT_RECORDS:-net.SA1.att.w1k,net.SA1.mlp.conv2,net.SA1.att.w2+softmax,net.SA1.att.v,net.FP1.mlp2.conv1} > gpurun_out/timeline_lat.txt 2>&1
timeout 300 python tools/tc_timeline.py pos 256 ${POS_RECORDS:-net.SA0.mlp.conv1,net.SA1.att.w1k,net.FP1.mlp2.conv1} > gpurun_out/timeline_pos.txt 2>&1
cat gpurun_out/timeline_lat.txt gpurun_out/timeline_pos.txt
