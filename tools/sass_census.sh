#!/bin/bash
# Instruction census of the shipped library: proves which tensor-core / TMA / TMEM instructions the kernels contain.
# usage: tools/sass_census.sh [lib] > profiles/r2_sass_census.txt        (runs without a GPU)
LIB=${1:-ml_conformer_generator_b200/libmlcg_b200.so}
echo "# cuobjdump -sass $LIB  (sm_100a), sources sha256 $(cat $LIB.stamp 2>/dev/null | cut -c1-16)"
cuobjdump -sass "$LIB" > /tmp/mlcg_all.sass
echo "## whole library"
for m in UTCHMMA UTCHMMA.2CTA UTCQMMA LDTM STTM UBLKCP UTMALDG UTCBAR "UTCBAR.2CTA.MULTICAST" UTCATOMSWS SYNCS MUFU.TANH MUFU.EX2 MUFU.RCP HMMA HGMMA; do
  printf "%-24s %6d\n" "$m" "$(grep -c -- " $m" /tmp/mlcg_all.sass)"
done
echo "## per kernel: UTC*MMA / LDTM / STTM / UBLKCP / MUFU.TANH"
awk '/Function :/ {name=$3} / UTC[A-Z]*MMA/ {m[name]++} / LDTM/ {l[name]++} / STTM/ {s[name]++} / UBLKCP/ {u[name]++} / MUFU.TANH/ {t[name]++} END {for (k in m) printf "%s %d %d %d %d %d\n", k, m[k], l[k], s[k], u[k], t[k]}' /tmp/mlcg_all.sass | c++filt | sort
