"""Stall samples per SASS instruction of the (first) kernel in an ncu report captured with --set full --import-source on:
the instructions with the most samples, each with its two main stall reasons and its executed count, then the headline
raw-page metrics.  usage: python profiles/sass_stalls.py <report.ncu-rep> [top N instructions]"""
import csv,sys,subprocess
rep=sys.argv[1]; n=int(sys.argv[2]) if len(sys.argv)>2 else 30
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
h=rows[hi]; ix={k:i for i,k in enumerate(h)}
data=rows[hi+1:]
tot=sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples',tot, 'instr', len(data))
agg={}
for r in data:
    for k in h:
        if k.startswith('stall_') and 'Not Issued' not in k:
            agg[k]=agg.get(k,0)+int(r[ix[k]] or 0)
print(sorted(agg.items(), key=lambda kv:-kv[1])[:8])
top=sorted(range(len(data)), key=lambda i:-int(data[i][ix['# Samples']] or 0))[:n]
for i in sorted(top):
    r=data[i]
    st={k:int(r[ix[k]] or 0) for k in h if k.startswith('stall_') and 'Not Issued' not in k}
    big=sorted(st.items(), key=lambda kv:-kv[1])[:2]
    print(i, r[ix['Source']][:64].ljust(64), r[ix['# Samples']], r[ix['Instructions Executed']], big)
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(out.splitlines())); h=r[0]; row=r[2]
for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active']:
    if k in h: print(k, row[h.index(k)], r[1][h.index(k)])
