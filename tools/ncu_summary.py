"""Condense one `ncu --set full` capture of the vote kernel into the JSON summary kept under profiles/.

usage: ncu_summary.py <report.ncu-rep> <votes_per_launch> <out.json> ["<command that produced the report>"]
The metric names are the ones SURVEY.md 8(d) lists; per-vote figures are derived from votes_per_launch
(printed by tools/profile_vote.py / bench.py for the same workload)."""
import csv, json, subprocess, sys

rep, votes, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
cmd = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
names, units, vals = rows[0], rows[1], rows[2]
want_prefix = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
               "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
               "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
               "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
               "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
               "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
               "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
               "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
               "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
               "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum",
               "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
               "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
               "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
               "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
               "smsp__average_warps_issue_stalled")
j = {}
for n, u, v in zip(names, units, vals):
    if n == "Kernel Name":
        j["kernel"] = v
    if any(n.startswith(p) for p in want_prefix):
        try:
            j[n] = {"value": float(v.replace(",", "")), "unit": u}
        except ValueError:
            pass
g = lambda k: j[k]["value"]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
dram = sum(g(k) * scale[j[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
j["votes_per_launch"] = votes
j["dram_bytes_per_launch"] = dram
j["dram_bytes_per_vote"] = dram / votes
j["thread_instructions_per_vote"] = g("smsp__inst_executed.sum") * 32 / votes
j["atoms_wavefronts_per_instruction"] = g("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum") / g("smsp__inst_executed_op_shared_atom.sum")
j["command"] = cmd
json.dump(j, open(out, "w"), indent=1)
print(json.dumps({k: j[k] for k in ("dram_bytes_per_vote", "thread_instructions_per_vote", "atoms_wavefronts_per_instruction")}))
