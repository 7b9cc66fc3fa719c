#!/bin/bash
for cs in 544 384 256 192; do
echo "== default lib, cam chunk $cs"; STBA_CAM_CHUNK=$cs timeout 200 python tools/time_lin.py | cut -c1-120
done
echo "== noalloc lib"; STBA_LIB=/root/repo/slam-tricks_b200/libstba_noalloc.so timeout 200 python tools/time_lin.py 10 | cut -c1-120
echo "== noalloc lib chunk 256"; STBA_CAM_CHUNK=256 STBA_LIB=/root/repo/slam-tricks_b200/libstba_noalloc.so timeout 200 python tools/time_lin.py | cut -c1-120
echo "== 10M chunk 512"; STBA_CAM_CHUNK=512 timeout 200 python tools/time_lin.py 10 | cut -c1-120
