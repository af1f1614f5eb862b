# ncu --set full of the attention cores: fwd2 (2 CTAs/SM) + split bwd, and the round-1 kernels, ViT shape
set -x
O=gpurun_out/r2ab; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 2 -o $O/attn_new_vit -f python tools/attn_one.py vit > /dev/null 2>&1
DC_ATTN_FWD_V1=1 DC_ATTN_BWD_V1=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 2 -o $O/attn_old_vit -f python tools/attn_one.py vit > /dev/null 2>&1
ls -la $O
