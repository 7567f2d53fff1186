for ks in 2 4 5; do echo "kshift $ks"; LRM_TC_KSHIFT=$ks bash tools/ab_var.sh product; done
echo "cell 2.5/640"; AB_CELL=2.5 AB_DIM=640 bash tools/ab_var.sh product
echo "cell 2/768"; AB_CELL=2 AB_DIM=768 bash tools/ab_var.sh product
echo "cell 4/384"; AB_CELL=4 AB_DIM=384 bash tools/ab_var.sh product
echo "bricks"; LRM_TC_BRICKS=1 bash tools/ab_var.sh product
