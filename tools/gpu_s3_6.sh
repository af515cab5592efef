mkdir -p gpurun_out
export SLIDE_B200_LIB=$PWD/slide_b200/libslide_b200_tl.so
timeout 300 python tools/tc_timeline.py lat 256 net.SA1.att.w1k,net.SA1.mlp.conv2,net.SA1.att.w2+softmax,net.SA1.mlp.conv1,net.SA1.att.v,net.SA0.mlp.conv1 > gpurun_out/tl_lat2.txt 2>&1
timeout 300 python tools/tc_timeline.py pos 256 net.SA0.mlp.conv1,net.SA1.att.w1k,net.SA1.mlp.conv2 > gpurun_out/tl_pos2.txt 2>&1
cat gpurun_out/tl_lat2.txt gpurun_out/tl_pos2.txt
