"""Aggregate an `ncu --page source --csv` dump of one kernel: instruction mix, stall mix, hottest SASS ranges."""
import sys
import pandas as pd
df = pd.read_csv(sys.argv[1], skiprows=1, dtype=str)
df = df[df['Address'].str.startswith('0x', na=False)].copy()
for c in ('Instructions Executed', '# Samples', 'Thread Instructions Executed'):
    df[c] = pd.to_numeric(df[c], errors='coerce').fillna(0)
df['op'] = df['Source'].str.strip().str.replace(r'^@!?U?P\d+\s+', '', regex=True).str.split().str[0]
df['opc'] = df['op'].str.split('.').str[0]
tot_inst = df['Instructions Executed'].sum(); tot_s = df['# Samples'].sum()
print('total warp inst %.4g samples %d sass lines %d  thread-inst/warp-inst %.1f' % (tot_inst, tot_s, len(df), df['Thread Instructions Executed'].sum() / tot_inst))
g = df.groupby('opc').agg(inst=('Instructions Executed', 'sum'), samp=('# Samples', 'sum'), n=('op', 'count')).sort_values('inst', ascending=False)
g['inst%'] = 100 * g.inst / tot_inst; g['samp%'] = 100 * g.samp / tot_s
print(g.head(int(sys.argv[2]) if len(sys.argv) > 2 else 24).round(2).to_string())
