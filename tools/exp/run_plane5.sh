#!/bin/bash
cd tools/exp/bin
echo "--- ~17M"
for b in pl2u pl2z; do ./$b 128 20 8; done
for b in pl3u pl3z pl3y pl3x; do ./$b 85 20 8; done
./pl4n 64 20 8
./pl1z 256 20 8
echo "--- 100M"
for b in pl2u pl2z pl3u pl3z pl3y pl3x pl4n pl1z; do ./$b 0 10 8 | grep mode; done
