#!/bin/bash
cd tools/exp/bin
echo "--- ~17M"
./pl4u 64 20 8; ./pl4z 64 20 8 | grep mode; ./pl2u 128 20 8; ./pl3u 85 20 8; ./pl3x 85 20 8 | grep mode; ./pl1a 256 20 8 | grep mode
echo "--- 100M"
for b in pl4u pl4z pl2u pl3u pl3x pl1a; do ./$b 0 10 8 | grep mode; done
