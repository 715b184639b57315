#!/bin/bash
# BASELINE.json configs[3]: throughput over the bundled same-sequencer (acid, q-score) model pairs, 100 bp synthetic
# reads drawn from each pair's own models, 4 M reads per pair, device-resident, both directions.
pairs=(
 "ERR174310__human__illumina_hiseq_2000__acids SRR2962693__human__illumina_hiseq_2500__q_scores"
 "SRR2962693__human__illumina_hiseq_2500__acids SRR2962693__human__illumina_hiseq_2500__q_scores"
 "SRR8861483__human__illumina_novaseq_6000__acids SRR8861483__human__illumina_novaseq_6000__q_scores"
 "SRR18908372__cat__illumina_novaseq_6000__acids SRR18908372__cat__illumina_novaseq_6000__q_scores"
 "SRR5373739__cat__illumina_hiseq_2500__acids SRR5373739__cat__illumina_hiseq_2500__q_scores"
 "m64187e__sars_cov_2__sequel_ii_e__acids m64187e__sars_cov_2__sequel_ii_e__q_scores"
 "ERR174310__human__illumina_hiseq_2000__acids SRR20210997__salmonella__illumina_hiseq_2500__q_scores"
 "SRR8861483__human__illumina_novaseq_6000__acids ERR5462922__ebov__illumina_iseq_100__q_scores"
)
echo "| acid model | q-score model | compress GB/s | decompress GB/s | B/read compat | native size | kernels |"
echo "|---|---|---|---|---|---|---|"
for p in "${pairs[@]}"; do
  set -- $p
  python bench.py --reads 4000000 --acid "$1" --q "$2" --steps 3 --no-e2e --no-cpu-baseline --no-fastq 2>/dev/null > /tmp/pair.json
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open("/tmp/pair.json")); o = d["other_mode"]
short = lambda n: n.split("__")[0] + " " + n.split("__")[-1]
static = "specialised" if d["roofline"]["kernels_ms_per_step"] else ""
print(f"| {short(sys.argv[1])} | {short(sys.argv[2])} | {d['compress_GBps']:.0f} | {d['decompress_GBps']:.0f} | {d['container_bytes_per_read']:.1f} | {o['size_vs_main_mode']:.3f} | enc {d['roofline']['kernels_ms_per_step']['encode']:.1f} ms, dec {d['roofline']['kernels_ms_per_step']['decode']:.1f} ms |")
PY
done
