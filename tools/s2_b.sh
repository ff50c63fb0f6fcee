# ncu of k_hp_pend / k_vpend at B=512 + digest on the box
mkdir -p gpurun_out
bash tools/ncu_one.sh "k_hp_pend|k_vpend" hpp
python tools/ncu_digest.py gpurun_out/hpp.ncu-rep 30 > gpurun_out/hpp_digest.txt 2>&1
