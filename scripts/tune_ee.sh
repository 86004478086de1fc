for hpl in 1 2; do for cw in 1 2 4 8; do
echo "== HPL_A=$hpl CW_A=$cw"; RSAC_EE_HPL_A=$hpl RSAC_EE_CW_A=$cw python scripts/quick_ee.py 1024 0 2>&1 | grep -A1 "last sweep" | tail -2
done; done
for hpl in 1 2; do for cw in 2 8; do
echo "== HPL_B=$hpl CW_B=$cw"; RSAC_EE_HPL_B=$hpl RSAC_EE_CW_B=$cw python scripts/quick_ee.py 1024 0 2>&1 | grep -A1 "last sweep" | tail -2
done; done
