"""Histogram of the coal-mining disaster dates used by the reference tutorial
(docs/src/advanced_tutorial_lit.jl:50-70: `read_coal_mining_data(..., 1)`), written as a small
fixture so that tests and the GPU box never read /root/reference at run time.

    python tests/golden/make_coal_mining.py        # needs /root/reference/docs/src/coal_mining_data.tsv

Restates read_coal_mining_data: intervals in days -> fractional years from the initial year ->
StatsBase `fit(Histogram, dates, left:binsize:right)` (left-closed bins, points beyond the last edge
dropped)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/docs/src/coal_mining_data.tsv"


def main():
    init_year, days = None, []
    for line in open(SRC):
        line = line.strip()
        if line.startswith("#"):
            if "init_date" in line:
                init_year = float(line[1:].strip().split("\t")[1].split("-")[0])
            continue
        if line:
            days.append(int(line))
    dates = init_year + np.cumsum(np.array(days, dtype=np.float64)) / 365.0
    binsize = 1.0
    left = dates[0]
    num_bins = int((dates[-1] - left) // binsize)
    edges = left + binsize * np.arange(num_bins + 1)
    counts = np.zeros(num_bins, dtype=np.int64)
    for d in dates:
        b = int(np.floor((d - left) / binsize))
        if 0 <= b < num_bins:
            counts[b] += 1
    np.savez_compressed(os.path.join(HERE, "coal_mining_counts.npz"), counts=counts, edges=edges, init_year=init_year)
    print("bins", num_bins, "events", int(counts.sum()), "of", len(days), "first", counts[:12])


if __name__ == "__main__":
    main()
