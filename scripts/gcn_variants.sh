for v in 1 3 5 7 0; do echo "gcn_snip=$v"; python scripts/gcn_only.py 2048 gcn_snip=$v; done
