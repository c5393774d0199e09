"""Summarise an .ncu-rep: key metrics per kernel + SASS hot-spot histogram. Usage: ncu_summary.py rep [kernel-regex]"""
import csv, subprocess, sys, io
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr,units=rows[0],rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
want+=[h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for w in want:
    if w in hdr:
        i=hdr.index(w); vals=[r[i][:28] for r in rows[2:]]
        if w.startswith('smsp__average_warps') and all(float(v or 0)<0.3 for v in vals): continue
        print(w.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''), units[i], vals)
if len(sys.argv)>2:
    src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+sys.argv[2]],capture_output=True,text=True).stdout
    rows=list(csv.reader(io.StringIO(src)))
    h=rows[1]; ie=h.index('Instructions Executed'); s=h.index('Source'); sm=h.index('# Samples'); at=h.index('Avg. Threads Executed')
    data=[(int(r[ie]),int(r[sm]),r[s].strip(), r[at]) for r in rows[2:] if len(r)>ie and r[ie].isdigit()]
    tot=sum(d[0] for d in data); ts=sum(d[1] for d in data) or 1
    print('total warp-inst',tot,'sass lines',len(data))
    step=int(sys.argv[3]) if len(sys.argv)>3 else 50
    for k in range(0,len(data),step):
        blk=data[k:k+step]; c=sum(d[0] for d in blk); sa=sum(d[1] for d in blk)
        ops=[d[2].split()[0] for d in blk]
        tags=[o for o in ops if o.startswith(('BAR','CALL','RET','ATOM','RED','SHFL','MUFU.RCP','LDG','STG','LDS','STS','BRA'))]
        from collections import Counter
        print(k,'inst %.1f%%'%(100*c/tot),'samp %.1f%%'%(100*sa/ts),'thr',blk[len(blk)//2][3],dict(Counter(tags)))
