import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith('{')][-1])  # (NCCL prints its version to stdout)
        print(path.split('/')[-1], 'n', d['n_gpus'], d['scaling'], 'ms/step %.4f' % d['ms_per_step'], 'value %.4e' % d['value'],
              'pair_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'],
              'opt_call_ms', d['config'].get('optimizer_call_ms_rank0'), '| e2e ms %.4f' % d['e2e']['ms_per_step'],
              'value %.4e' % d['e2e']['value'], 'h2d', d['e2e']['h2d_bytes_per_step'], '| nvlink', d['config'].get('nvlink_rank0'),
              '| launches', d['gpu_launches'], 'clocks', d.get('clocks'))
    except Exception as e:  # noqa: BLE001
        print(path, 'unreadable:', e)
