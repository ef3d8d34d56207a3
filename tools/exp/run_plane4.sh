#!/bin/bash
cd tools/exp/bin
echo "--- 116 cells"
for b in pl4g pl4gs pl4gt; do for c in 1 2 4 8; do ./$b 0 10 $c | grep "mode="; done; done
for b in pl4c pl4cs pl4ct; do for c in 1 2 4 7; do ./$b 0 10 $c | grep "mode="; done; done
echo "--- 64 cells"
for b in pl4g pl4gs pl4gt pl4c pl4cs pl4ct; do for c in 1 2 3 4 6 8 16; do ./$b 64 20 $c | grep "mode="; done; done
