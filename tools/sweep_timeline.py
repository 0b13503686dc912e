"""Summary of a sweep stage timeline (PCM_STAGE_TIMELINE=<path> python bench.py --workload sweep, pcm/stages.py):
when the clip-preparation steps ran and on which thread, how long the fits / sequences took, and how many fitting and
sequence threads were busy per quarter second -- the critical path of the grid sweep, which the per-stage sums cannot
show.  Usage: python tools/sweep_timeline.py timeline.json"""
import json, collections, sys
ev = json.load(open(sys.argv[1]))
print(sys.argv[1], len(ev), "events; end %.2f" % max(e["end"] for e in ev))
for st in ["prepare_clips"]:
    main = [e for e in ev if e["stage"]==st][0]["thread"]
for e in sorted([e for e in ev if e["stage"] in ("tracker_boxes","quickshift_maps","felzenszwalb_maps","sift_detect")], key=lambda e:e["start"]):
    if e["end"]-e["start"]>0.02: print("  t%-2d %-20s %.3f -> %.3f (%.3f)" % (e["thread"], e["stage"], e["start"], e["end"], e["end"]-e["start"]))
def hist(stage):
    xs=[e for e in ev if e["stage"]==stage]
    if xs: print("%-14s %4d first start %.2f last end %.2f; mean dur %.3f max %.3f sum %.1f" % (stage, len(xs), min(e["start"] for e in xs), max(e["end"] for e in xs), sum(e["end"]-e["start"] for e in xs)/len(xs), max(e["end"]-e["start"] for e in xs), sum(e["end"]-e["start"] for e in xs)))
for s in ["prefit","fit_forest","fit_pca","train_rows","export_model","sequence","enqueue","gpu_wait"]:
    hist(s)
def occ(stage):
    bins=collections.Counter()
    for e in ev:
        if e["stage"]==stage:
            t=e["start"]
            while t<e["end"]:
                b=int(t/0.25); nx=(b+1)*0.25
                bins[b]+=min(e["end"],nx)-t; t=nx
    print(stage, "occupancy per 0.25 s:", [round(bins[k]/0.25,1) for k in range(int(max(e["end"] for e in ev)/0.25)+1)])
occ("sequence"); occ("prefit")
