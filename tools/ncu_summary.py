import sys, csv, subprocess
rep = sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
pats = sys.argv[2:] or ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct','sm__warps_active.avg.pct','launch__registers_per_thread','bank_conflicts','wavefronts_mem_shared.sum','sm__inst_executed.sum','sm__throughput.avg.pct','l1tex__throughput.avg.pct','lts__throughput.avg.pct','issue_active.avg.pct','occupancy_limit','warp_issue_stalled.*per_warp_active','lts__t_sector_hit_rate','l1tex__t_sector_hit_rate','launch__shared_mem_per_block','achieved_occupancy','sm__pipe_fma_cycles_active','sm__inst_executed_pipe_lsu','lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
import re
for r in rows[2:]:
    print('----')
    for i,h in enumerate(hdr):
        if any(re.search(p,h) for p in pats):
            v = r[i]
            if v in ('0','','n/a'): continue
            print(f"{h} = {v} {rows[1][i]}")
