"""Summarise `ncu --page raw --csv` and `--page source --csv` exports (run here, no GPU needed).
   python tools/ncu_summary.py raw gpurun_out/x_raw.csv            # one block of key metrics per kernel
   python tools/ncu_summary.py src gpurun_out/x.ncu-rep <kernel-regex>   # instruction mix, stall reasons, hot lines"""
import collections, csv, re, subprocess, sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu.sum', 'smsp__inst_executed_op_shared_st.sum', 'lts__t_sector_hit_rate.pct']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('---', r[idx['Kernel Name']][:110])
        for w in WANT:
            if w in idx:
                print(f'   {w:70s} {r[idx[w]]:>18s} {units[idx[w]]}')


def src(rep, kernel, top='18'):
    """kernel: substring of the demangled name (template arguments included, e.g. 'gemm_tc_kernel<(int)1') or the index of
    the kernel in the report. `ncu --kernel-name regex:` only sees the base name, so the page is split here."""
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            if not blocks or blocks[-1]['name'] != cur['name'] or blocks[-1]['rows']:
                blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    blocks = [b for i, b in enumerate(blocks) if i % 2 == 0] if len(blocks) > 1 and blocks[0]['name'] == blocks[1]['name'] else blocks
    if kernel.isdigit():
        b = blocks[int(kernel)]
    else:
        cand = [b for b in blocks if kernel in b['name']]
        if not cand:
            print('kernels in the report:'); [print(f'  [{i}]', b['name'][:120]) for i, b in enumerate(blocks)]
            return
        b = cand[0]
    top = int(top)
    hdr = b['rows'][0]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in b['rows'][1:] if len(r) > 8]

    def num(x):
        try: return float(x)
        except Exception: return 0.0
    ti = sum(num(r[idx['Instructions Executed']]) for r in data)
    ts = sum(num(r[idx['# Samples']]) for r in data)
    print(b['name'][:120])
    print(f'{len(data)} SASS lines, {ti/1e6:.1f} M warp-instructions, {ts:.0f} samples')
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[idx['Source']])
        o = m.group(2) if m else r[idx['Source']][:16]
        o = '.'.join(o.split('.')[:2])
        op[o] += num(r[idx['Instructions Executed']]); ops[o] += num(r[idx['# Samples']])
    for o, c in op.most_common(24):
        print(f'  {o:16s} {c/1e6:9.1f} M {100*c/ti:5.1f}%   samples {100*ops[o]/max(ts,1):5.1f}%')
    st = {h: sum(num(r[idx[h]]) for r in data) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h}
    s = sum(st.values()) or 1
    print('  stalls:', {k[6:]: round(100 * v / s, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / s > 0.01})
    for r in sorted(data, key=lambda r: -num(r[idx['# Samples']]))[:top]:
        print(f"  {r[idx['# Samples']]:>7s} smp {r[idx['Instructions Executed']]:>10s} x  {r[idx['Source']].strip()[:80]}")


if __name__ == '__main__':
    (raw if sys.argv[1] == 'raw' else src)(*sys.argv[2:])
