mkdir -p gpurun_out
export EKFSLAM_DEFER_HI=0
echo "2 CTAs/SM"; python tools/split_probe.py 4096 1 2
echo "1 CTA/SM"; EKFSLAM_DD_CTAS_PER_SM=1 python tools/split_probe.py 4096 1 2 4
