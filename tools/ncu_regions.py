"""Executed instructions / stall samples of one kernel of an .ncu-rep, grouped into source regions.

    python tools/ncu_regions.py <report.ncu-rep> <kernel substring> <voxels per launch>

The line ranges below describe the source as it stood when profiles/r02m_* were captured (commit a9463eb: the ELBO
kernels were still inside elbo.cu).  Since then the kernel bodies moved to *_kernels.cuh; re-derive the ranges before
using this on a new capture."""
import csv, io, subprocess, sys, collections
rep=sys.argv[1]; sub=sys.argv[2]; nvox=float(sys.argv[3])
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
fn=fpath=hdr=None; seen=set(); agg=collections.Counter(); smp=collections.Counter()
def region(f, line):
    if f=='bessel.cuh': return 'quadrature: bessel'
    if f=='qbold_core.cuh':
        if 360<=line<=540: return 'quadrature: sched loop/flush/reduce'
        if line in range(240,262): return 'work counter'
        if 160<=line<=200: return 'voxel_phys'
        if 205<=line<=250: return 'tau_signal'
        if 262<=line<=275: return 'node0'
        return 'core other %d'%line
    if f=='rng.cuh':
        if 25<=line<=46: return 'philox + u01'
        if 48<=line<=62: return 'accurate box-muller (reparam)'
        return 'mc box-muller'
    if f=='elbo.cu':
        if 44<=line<=92: return 'load_dists'
        if 94<=line<=106 or line==27: return 'draw/sigmoid'
        if 108<=line<=117: return 'kl: mvn_nll'
        if 119<=line<=153: return 'cold paths'
        if 155<=line<=260: return 'kl: loop/accum/reduce'
        if 460<=line<=520: return 'k_elbo_pair: prologue/loads'
        if 521<=line<=545: return 'k_elbo_pair: quadrature glue'
        if 546<=line<=595: return 'k_elbo_pair: NLL + grads'
        if 596<=line<=640: return 'k_elbo_pair: KL glue + stores'
        return 'elbo other %d'%line
    return f
for r in rows:
    if not r: continue
    if r[0]=='File Path': fpath=r[1]
    elif r[0]=='Function Name': fn=r[1]
    elif r[0]=='Line No': hdr=r
    elif hdr and r[0].isdigit() and fn and sub in fn:
        key=(fpath,int(r[0]))
        if key in seen: continue
        seen.add(key)
        def num(name):
            v=r[hdr.index(name)-len(hdr)].replace(',','')
            return float(v) if v not in('-','') else 0.0
        reg=region(fpath.split('/')[-1],int(r[0]))
        agg[reg]+=num('Instructions Executed'); smp[reg]+=num('# Samples')
tot=sum(agg.values()); ts=sum(smp.values())
for k,v in agg.most_common():
    print('%-45s %6.1f%% inst  %6.1f%% samples' % (k, 100*v/tot, 100*smp[k]/ts))
