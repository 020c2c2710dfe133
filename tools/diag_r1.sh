set -x
python tools/tune.py C2 BLOCKS=2,3,4 2>&1 | tail -3
python tools/tune.py C2 ORDER=0,1 CHUNK=16,32 2>&1 | tail -4
python tools/tune.py C2 EXIT=12,16,20,24 MAX_HK=1,2 2>&1 | tail -8
python tools/tune.py C4 BLOCKS=2,3 ORDER=0,1 2>&1 | tail -4
