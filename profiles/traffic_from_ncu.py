"""DRAM traffic per launch of the captured kernels (one `ncu --set full` capture each, profiles/capture_r02.sh) next to their
algorithmic bytes -> profiles/r02_traffic.json (bench.py's roofline.traffic reads the dominant kernel's row from it).

    python profiles/traffic_from_ncu.py gpurun_out r02a r02b
"""
import csv
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
E = 16 * 512 * 512 * 64 * 2          # one 16 x 512 x 512 x 64 bf16 activation
# capture name -> (what it ran, algorithmic bytes = every live tensor read once + written once, algorithmic FLOPs)
CASES = {
    "halo_wgrad_l0": ("conv0_0.conv2 weight gradient, 16x512x512, 64->64, 3x3", 2 * E + 64 * 64 * 9 * 4, 2.0 * 16 * 512 * 512 * 64 * 64 * 9),
    "halo_wgrad_l1c": ("conv1_1.conv1 weight gradient, 16x256x256, 384->128, 3x3", 16 * 256 * 256 * (384 + 128) * 2 + 384 * 128 * 9 * 4,
                       2.0 * 16 * 256 * 256 * 384 * 128 * 9),
    "thin_wgrad_gb": ("SPADE gamma|beta weight gradient, 16x512x512, 8->128, 3x3", 16 * 512 * 512 * (8 + 128) * 2 + 8 * 128 * 9 * 4,
                      2.0 * 16 * 512 * 512 * 4 * 128 * 9),
    "s2_dgrad_l0": ("D block1 data gradient, stride 2, 16x512x512 <- 16x256x256, 64<-64", E + E // 4, 2.0 * 16 * 256 * 256 * 64 * 64 * 9),
    "s2_fwd_l0": ("D block1 forward, stride 2, 16x512x512 -> 16x256x256, 64->64", E + E // 4, 2.0 * 16 * 256 * 256 * 64 * 64 * 9),
    "s2_wgrad_l0": ("D block1 weight gradient, stride 2", E + E // 4 + 64 * 64 * 9 * 4, 2.0 * 16 * 256 * 256 * 64 * 64 * 9),
    "halo_fwd_l0": ("conv0_0.conv2 forward, 16x512x512, 64->64, 3x3", 2 * E, 2.0 * 16 * 512 * 512 * 64 * 64 * 9),
    "halo_dgrad_l0": ("conv0_0.conv2 data gradient, 16x512x512, 64<-64, 3x3", 2 * E, 2.0 * 16 * 512 * 512 * 64 * 64 * 9),
    "bn_bwd_apply": ("BN backward apply (dy, y, x -> dx), 16x512x512x64", 4 * E, 0.0),
    "bn_bwd_reduce": ("BN backward reduce (dy, y, x -> sums), 16x512x512x64", 3 * E, 0.0),
    # second batch of captures (profiles/capture_r02b.sh): the kernels changed in round 2
    "halo_wgrad128_l1c": ("conv1_1.conv1 weight gradient, 128-wide co tiles in two tap-pair passes", 16 * 256 * 256 * (384 + 128) * 2 + 384 * 128 * 9 * 4,
                          2.0 * 16 * 256 * 256 * 384 * 128 * 9),
    "halo_wgrad64_l0c": ("conv0_1.conv1 weight gradient (192 -> 64 over the virtual concat), one wave of CTAs", 16 * 512 * 512 * (192 + 64) * 2 + 192 * 64 * 9 * 4,
                         2.0 * 16 * 512 * 512 * 192 * 64 * 9),
    "halo_wgrad64_l0": ("conv0_0.conv2 weight gradient, two issuing warps", 2 * E + 64 * 64 * 9 * 4, 2.0 * 16 * 512 * 512 * 64 * 64 * 9),
    "s2_dgrad_merged_l0": ("D block1 data gradient, stride 2, four parity classes in one CTA", E + E // 4, 2.0 * 16 * 256 * 256 * 64 * 64 * 9),
    "halo_fwd_l0_2issuers": ("conv0_0.conv2 forward, two issuing warps", 2 * E, 2.0 * 16 * 512 * 512 * 64 * 64 * 9),
    "halo_thin_x2map_fwd": ("SPADE x2map forward 64 -> 3 (stored 8), 16x512x512", E + E // 8, 2.0 * 16 * 512 * 512 * 64 * 3 * 9),
    # third batch (profiles/capture_r02d.sh, capture_r02e.sh): thin-input forward before / after the straight-line issue path and the
    # second epilogue group, the level-0 forward after it, the halo-form stride-2 data gradient
    "halo_thin_in_dconv0_fwd": ("D block0 forward 3 (stored 8) -> 64, 16x512x512, generic issue loop, one epilogue group", E + E // 8,
                                2.0 * 16 * 512 * 512 * 3 * 64 * 9),
    "halo_thin_in_gb_fwd": ("SPADE gamma|beta forward 4 (stored 8) -> 128, 16x512x512, generic issue loop, one epilogue group", 2 * E + E // 8,
                            2.0 * 16 * 512 * 512 * 4 * 128 * 9),
    "halo_thin_in_dconv0_fwd_lean": ("D block0 forward 3 (stored 8) -> 64, straight-line issue + two epilogue groups", E + E // 8,
                                     2.0 * 16 * 512 * 512 * 3 * 64 * 9),
    "halo_fwd_l0_lean": ("conv0_0.conv2 forward, straight-line issue path", 2 * E, 2.0 * 16 * 512 * 512 * 64 * 64 * 9),
    "s2_dgrad_halo_l0": ("D block1 data gradient, stride 2, halo formulation (one dy box, four class accumulators)", E + E // 4,
                         2.0 * 16 * 256 * 256 * 64 * 64 * 9),
}


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, r = rows[0], rows[1], rows[2]

    def val(k):
        i = h.index(k)
        return float(r[i].replace(",", "")) * UNIT.get(u[i], 1.0)
    return {"kernel": r[h.index("Kernel Name")][:80], "dram_read_bytes": val("dram__bytes_read.sum"),
            "dram_write_bytes": val("dram__bytes_write.sum"), "duration_us": val("gpu__time_duration.sum"),
            "tensor_pipe_active_pct": float(r[h.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")])}


def main():
    d, tags = sys.argv[1], sys.argv[2:]
    res = []
    for name, (what, alg, flops) in CASES.items():
        hits = [(t, os.path.join(d, "%s_%s.ncu-rep" % (t, name))) for t in tags]
        hits = [(t, r) for t, r in hits if os.path.exists(r)]
        if not hits:
            continue
        tag, rep = hits[0]
        m = read(rep)
        m.update({"capture": "%s_%s" % (tag, name), "what": what, "algorithmic_bytes": alg,
                  "dram_bytes": m["dram_read_bytes"] + m["dram_write_bytes"]})
        m["traffic_over_algorithmic"] = round(m["dram_bytes"] / alg, 3)
        if flops:
            m["tflops_under_ncu"] = round(flops / (m["duration_us"] * 1e-6) / 1e12, 1)
        else:
            m["GBps_under_ncu"] = round(alg / (m["duration_us"] * 1e-6) / 1e9, 1)
        res.append(m)
        print("%-16s %-60s dram %.3f GB (x%.2f of algorithmic) %.1f us tensor %.1f%%" % (
            name, what[:60], m["dram_bytes"] / 1e9, m["traffic_over_algorithmic"], m["duration_us"], m["tensor_pipe_active_pct"]))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_traffic.json"), "w") as f:
        json.dump({"how": "one `ncu --set full --clock-control none` capture per kernel (profiles/capture_r02.sh), dram__bytes_read.sum + "
                          "dram__bytes_write.sum per launch; durations under ncu are cold-cache and serialised", "kernels": res}, f, indent=1)


if __name__ == "__main__":
    main()
