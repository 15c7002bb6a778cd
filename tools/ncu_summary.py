"""Summarise ncu outputs brought back in gpurun_out/: launch list shares, selected raw metrics, SASS opcode mix."""
import collections, csv, io, re, subprocess, sys

def launches(path, top=14):
    top = int(top)
    rows=[r for r in csv.reader(open(path)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    agg=collections.OrderedDict()
    for r in rows[1:]:
        try: v=float(r[vi].replace(',',''))
        except: continue
        a=agg.setdefault(r[ki],[0,0.0]); a[0]+=1; a[1]+=v
    tot=sum(a[1] for a in agg.values())
    out=["| share | total us | launches | us/launch | kernel |","|---|---|---|---|---|"]
    for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:top]:
        out.append(f"| {100*a[1]/tot:.1f}% | {a[1]/1e3:.1f} | {a[0]} | {a[1]/1e3/a[0]:.1f} | `{k[:80]}` |")
    return "\n".join(out)

WANT=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__thread_inst_executed_per_inst_executed.ratio']

def raw(rep):
    txt=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
    rr=list(csv.reader(io.StringIO(txt))); h=rr[0]
    names=[r[h.index('Kernel Name')][:48] for r in rr[2:]]
    out=["| metric | "+" | ".join(names)+" |","|---|"+"---|"*len(names)]
    for w in WANT:
        if w in h:
            i=h.index(w); out.append(f"| {w} [{rr[1][i]}] | "+" | ".join(r[i] for r in rr[2:])+" |")
    return "\n".join(out)

def opmix(rep, kernel_idx=None, top=22):
    txt=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
    blocks=txt.split('"Kernel Name",')[1:]
    res=[]
    for b in blocks:
        rows=list(csv.reader(io.StringIO('"Kernel Name",'+b)))
        name=rows[0][1]; h=rows[1]; ia=h.index('Source'); ie=h.index('Instructions Executed'); isamp=h.index('# Samples')
        ops=collections.Counter(); samp=collections.Counter(); tot=tots=0
        for r in rows[2:]:
            if len(r)<=ie: continue
            try: n=int(r[ie]); s=int(r[isamp])
            except: continue
            m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia]); op=m.group(2).split('.')[0] if m else '?'
            ops[op]+=n; samp[op]+=s; tot+=n; tots+=s
        out=[f"kernel `{name[:70]}`: {tot} warp instructions, {tots} samples","| opcode | executed | share | stall-sample share |","|---|---|---|---|"]
        for op,n in ops.most_common(top): out.append(f"| {op} | {n} | {100*n/tot:.1f}% | {100*samp[op]/max(1,tots):.1f}% |")
        res.append("\n".join(out))
    return "\n\n".join(res)

if __name__=="__main__":
    cmd=sys.argv[1]
    print({"launches":launches,"raw":raw,"opmix":opmix}[cmd](*sys.argv[2:]))
