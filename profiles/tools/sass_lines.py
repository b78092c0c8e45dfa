"""Attribute the per-instruction counters of an `ncu --set full --import-source on` capture to source lines.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > k.csv
    nvdisasm -gi -c file.cubin > all.sass          (cubin from `cuobjdump -xelf all lib.so`)
    python sass_lines.py k.csv all.sass <mangled-function-substring> <file.cu> <first-line> <last-line>

Instructions that come from inlined helpers (lines outside [first, last] or other files) are booked on the most recent
kernel-level line, so the table reads as "cost of each statement of the kernel body"."""
import csv, re, sys, collections

def main():
    kcsv, sass, fn, cu, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
    lines_of = []                      # per instruction (in order): (own line key, kernel-level line)
    cur_file, cur_line, klevel, inside, fresh = None, None, None, False, True
    for ln in open(sass):
        if ln.startswith(".text."):
            inside = fn in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            # nvdisasm -gi prints the inline chain innermost first; the first entry after an instruction starts a new chain
            if fresh:
                cur_file, cur_line = m.group(1), int(m.group(2))
                fresh = False
            f2, l2 = m.group(1), int(m.group(2))
            if f2.endswith(cu) and lo <= l2 <= hi:
                klevel = l2                                 # the last (outermost) match of the chain wins
            continue
        if re.match(r"\s*/\*[0-9a-f]+\*/", ln):
            fresh = True
            lines_of.append(((cur_file.split("/")[-1] if cur_file else "?", cur_line), klevel))
    rows = list(csv.reader(open(kcsv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break                                       # a second launch of the same kernel follows
        body.append(r)
    assert len(body) == len(lines_of), (len(body), len(lines_of))
    agg = collections.defaultdict(lambda: collections.Counter())
    own = collections.defaultdict(lambda: collections.Counter())
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    for r, (o, k) in zip(body, lines_of):
        inst = int(r[col["Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        op = r[col["Source"]].split()[0] if not r[col["Source"]].strip().startswith("@") else r[col["Source"]].split()[1]
        for key, d in ((k, agg), (o, own)):
            d[key]["inst"] += inst
            d[key]["samples"] += smp
            if op.startswith("HMMA"):
                d[key]["hmma"] += inst
            for s in stalls:
                d[key][s] += int(r[col[s]] or 0)
        tot["inst"] += inst
        tot["samples"] += smp
    print("total inst %d samples %d" % (tot["inst"], tot["samples"]))
    def show(d, title, n):
        print("==", title)
        for key, c in sorted(d.items(), key=lambda kv: -kv[1]["samples"])[:n]:
            top = sorted(((c[s], s[6:]) for s in stalls), reverse=True)[:3]
            print("%-26s inst %5.1f%%  hmma %5.1f%%  samples %5.1f%%  %s" % (
                str(key), 100.0 * c["inst"] / tot["inst"], 100.0 * c["hmma"] / max(1, tot["inst"]), 100.0 * c["samples"] / tot["samples"],
                " ".join("%s=%.0f%%" % (s, 100.0 * v / max(1, c["samples"])) for v, s in top)))
    show(agg, "by kernel-level line", int(sys.argv[7]) if len(sys.argv) > 7 else 45)
    show(own, "by own line (helpers)", 25)

main()
