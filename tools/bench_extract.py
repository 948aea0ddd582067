import sys, json, fileinput
# one-line view of bench.py JSON lines: files given as arguments, or stdin ('-' / no argument)
for line in fileinput.input():
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    out={k:(d[k].get('value') if isinstance(d.get(k),dict) else d.get(k)) for k in ['value','constrained','config3','config3_constrained','ipm_only','config4']}
    out['e2e']=d['e2e']['value']; out['ms']=d['ms_per_step']
    print(json.dumps(out))
