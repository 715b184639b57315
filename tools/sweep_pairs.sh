#!/bin/bash
# BASELINE.json configs[3] / SURVEY.md 8d config 4: throughput over the bundled same-sequencer (acid, q-score) model pairs
# (the 11 pairs of tests/test_gpu_variants.py), 100 bp synthetic reads drawn from each pair's own models, 4 M reads per pair,
# device-resident, both directions; then per-read selection among the 4 + 4 models quality 7 retains, both formats.
pairs=(
 "ERR174310__human__illumina_hiseq_2000__acids SRR2962693__human__illumina_hiseq_2500__q_scores SP0"
 "SRR2962693__human__illumina_hiseq_2500__acids SRR2962693__human__illumina_hiseq_2500__q_scores SP0"
 "SRR19549058__b_stabilis__illumina_hiseq_2500__acids SRR19549058__b_stabilis__illumina_hiseq_2500__q_scores SP0"
 "SRR8861483__human__illumina_novaseq_6000__acids SRR8861483__human__illumina_novaseq_6000__q_scores SP1"
 "m64187e__sars_cov_2__sequel_ii_e__acids m64187e__sars_cov_2__sequel_ii_e__q_scores SP2"
 "SRR18908372__cat__illumina_novaseq_6000__acids SRR18908372__cat__illumina_novaseq_6000__q_scores SP3"
 "SRR5373739__cat__illumina_hiseq_2500__acids SRR5373739__cat__illumina_hiseq_2500__q_scores SP4"
 "ERR5462922__ebov__illumina_iseq_100__acids ERR5462922__ebov__illumina_iseq_100__q_scores SP5"
 "SRR16141966__e_coli__illumina_hiseq_2500__acids SRR16141966__e_coli__illumina_hiseq_2500__q_scores SP6"
 "SRR19609907__pear__illumina_hiseq_2500__acids SRR19609907__pear__illumina_hiseq_2500__q_scores SP7"
 "SRR20210997__salmonella__illumina_hiseq_2500__acids SRR20210997__salmonella__illumina_hiseq_2500__q_scores generic"
)
echo "| acid model | q-score model | kernels | compress GB/s | decompress GB/s | B/read compat | native size | encode / decode ms |"
echo "|---|---|---|---|---|---|---|---|"
for p in "${pairs[@]}"; do
  set -- $p
  python bench.py --reads 4000000 --acid "$1" --q "$2" --steps 3 --no-e2e --no-cpu-baseline --no-fastq --no-extra-workloads 2>/dev/null > /tmp/pair.json
  python - "$1" "$2" "$3" <<'PY'
import json, sys
d = json.load(open("/tmp/pair.json")); o = d["other_mode"]
short = lambda n: n.split("__")[0] + " " + n.split("__")[-1]
k = d["roofline"]["kernels_ms_per_step"]
print(f"| {short(sys.argv[1])} | {short(sys.argv[2])} | {sys.argv[3]} | {d['compress_GBps']:.0f} | {d['decompress_GBps']:.0f} | {d['container_bytes_per_read']:.1f} | {o['size_vs_main_mode']:.3f} | {k['encode']:.2f} / {k['decode']:.2f} |")
PY
done
for mode in compat native; do
  python bench.py --workload hiseq100_select4 --mode $mode --reads 4000000 --steps 3 --no-e2e --no-cpu-baseline --no-fastq --no-extra-workloads --no-other-mode 2>/dev/null > /tmp/pair.json
  python - $mode <<'PY'
import json, sys
d = json.load(open("/tmp/pair.json"))
k = d["roofline"]["kernels_ms_per_step"]
enc = k.get("encode", k.get("encode_lane")); dec = k.get("decode", k.get("decode_lane"))
print(f"| 4 acid models | 4 q-score models (per-read selection, {sys.argv[1]}) | buckets | {d['compress_GBps']:.0f} | {d['decompress_GBps']:.0f} | {d['container_bytes_per_read']:.1f} | | {enc:.2f} / {dec:.2f} (score {k['score']:.2f}) |")
PY
done
