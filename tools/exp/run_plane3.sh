#!/bin/bash
cd tools/exp/bin
for b in pl*; do [ -x ./$b ] && timeout 120 ./$b 0 10; done
echo "--- chunk 8"
for b in pl4e pl4f pl4g pl3c pl3d; do ./$b 0 10 8 | grep "mode="; done
