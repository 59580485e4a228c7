set -x
python scripts/sweep.py D/8 "sell:R=512;G=2;U=4,sell:R=256;G=2;U=4,sell:R=128;G=2;U=4,sell:R=256;G=2;U=2,sell:R=128;G=2;U=2,sell:R=64;G=2;U=2,sell:R=128;G=1;U=2,sell:R=256;G=2;U=6,vector" 30 2>&1 | tee gpurun_out/sweep11.txt
python scripts/sweep.py crsmat170 "sell:R=512;G=2;U=4,sell:R=512;G=2;U=2,sell:R=256;G=2;U=2,sell:R=128;G=2;U=2,sell:R=512;G=1;U=2,vector" 50 2>&1 | tee -a gpurun_out/sweep11.txt
python scripts/sweep.py pl22 "sell:R=512;G=2;U=4,sell:R=512;G=2;U=2,sell:R=128;G=2;U=2,sell:R=512;G=2;U=2;C=32,sell:R=512;G=2;U=2;C=256" 30 2>&1 | tee -a gpurun_out/sweep11.txt
python scripts/sweep.py C "sell:R=512;G=2;U=4,sell:R=128;G=2;U=2,sell:R=256;G=2;U=2" 100 2>&1 | tee -a gpurun_out/sweep11.txt
python scripts/sweep.py A "sell:R=512;G=2;U=4,sell:R=64;G=1;U=2,sell:R=64;G=2;U=2,sell:R=128;G=2;U=2" 200 2>&1 | tee -a gpurun_out/sweep11.txt
