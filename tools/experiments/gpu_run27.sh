for pdl in 0 1; do for pair in 0 1; do
echo "pdl $pdl pair $pair"; B2D_PDL=$pdl B2D_PAIR=$pair timeout 300 python tools/diag.py time --batch 64 2>&1 | tail -2 | head -1
done; done
