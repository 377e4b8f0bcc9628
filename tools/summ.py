import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f,"ERR",e); continue
    r=d["roofline"]; n=d["steps"]*d["config"]["frames_per_step_per_gpu"]
    print(f.split("/")[-1], round(d["value"]), "fps", round(d["us_per_tile_frame"],2), "us/frame e2e", round(d["e2e"]["value"]),
          "us/frame by kernel", {k:round(v*1e3/n,2) for k,v in r["kernel_ms"].items()}, "dom", r["kernel"][:16], "frac", round(r["frac"],3), "clk", d["clocks"]["sm_mhz"])
