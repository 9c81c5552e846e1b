#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "default:"; python -m xmem2_b200.util.conv_bench "fuser 3x3 512->512" | head -1
echo "forced 128,2,6:"; XMEM_CONV_FORCE=128,2,6 python -m xmem2_b200.util.conv_bench "fuser 3x3 512->512" | head -1
echo "l3 1x1 1024->256 default / forced 128,2,6 / forced 64,3,6:"; python -m xmem2_b200.util.conv_bench "l3 1x1 1024->256" | head -1; XMEM_CONV_FORCE=128,2,6 python -m xmem2_b200.util.conv_bench "l3 1x1 1024->256" | head -1; XMEM_CONV_FORCE=64,3,6 python -m xmem2_b200.util.conv_bench "l3 1x1 1024->256" | head -1
echo "keyproj default / forced 64,2,6:"; python -m xmem2_b200.util.conv_bench "keyproj" | head -1; XMEM_CONV_FORCE=64,2,6 python -m xmem2_b200.util.conv_bench "keyproj" | head -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
