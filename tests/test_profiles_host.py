"""The committed ncu traffic summary carries the kernel families bench.py looks up for ``roofline.traffic`` (a renamed kernel would
silently turn the field into null)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_summary_has_the_families_bench_reads():
    import bench
    fam = bench.load_traffic()
    assert fam is not None, "profiles/r02_traffic.json (or r01) missing"
    gemm = fam.get("gemm_bf16_tcgen05_kernel")
    att = fam.get("attention_hpw_kernel") or fam.get("attention_kernel")
    assert gemm and gemm["launches"] > 1000 and 5e7 < gemm["dram_bytes_per_launch"] < 5e8
    assert att and att["launches"] >= 300 and 5e8 < att["dram_bytes_per_launch"] < 1.2e9      # ~ 812 MB algorithmic per decode-attention launch


def test_family_grouping_of_the_summariser(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("ncu_summarize", os.path.join(ROOT, "scripts", "ncu_summarize.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = ['"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"']
    def add(i, name, metric, unit, value):
        rows.append(f'"{i}","1","python","h","{name}","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","{metric}","{unit}","{value}"')
    for i, name in enumerate(["void lmrl::gemm_bf16_tcgen05_2cta_kernel<(int)1>(CUtensorMap_st)", "void lmrl::gemm_bf16_tcgen05_kernel<(int)0, (int)256, (bool)0>(CUtensorMap_st)",
                              "void lmrl::attention_hpw_kernel<(int)4>(const __nv_bfloat16 *)"]):
        add(i, name, "gpu__time_duration.sum", "us", "100")
        add(i, name, "dram__bytes_read.sum", "Mbyte", "10")
        add(i, name, "dram__bytes_write.sum", "Mbyte", "5")
    src = tmp_path / "launches.csv"
    src.write_text("\n".join(rows) + "\n")
    out = tmp_path / "traffic.json"
    mod.traffic(str(src), str(out))
    fam = json.load(open(out))["families"]
    assert fam["gemm_bf16_tcgen05_kernel"]["launches"] == 2 and abs(fam["gemm_bf16_tcgen05_kernel"]["dram_bytes_per_launch"] - 15e6) < 1
    assert fam["attention_hpw_kernel"]["launches"] == 1
