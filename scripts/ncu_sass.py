"""Aggregate an `ncu --page source --csv` dump of one kernel: instruction mix, stall mix, hottest SASS ranges."""
import sys
import pandas as pd
df = pd.read_csv(sys.argv[1], skiprows=1)
df['op'] = df['Source'].str.strip().str.replace(r'^@!?U?P\d+\s+', '', regex=True).str.split().str[0]
df['opc'] = df['op'].str.split('.').str[0]
tot_inst = df['Instructions Executed'].sum(); tot_s = df['# Samples'].sum()
print('total warp inst', tot_inst, 'samples', tot_s, 'sass lines', len(df))
g = df.groupby('opc').agg(inst=('Instructions Executed', 'sum'), samp=('# Samples', 'sum'), n=('op', 'count')).sort_values('inst', ascending=False)
g['inst%'] = 100 * g.inst / tot_inst; g['samp%'] = 100 * g.samp / tot_s
print(g.head(28).to_string())
stall_cols = [c for c in df.columns if c.startswith('stall_') and 'Not Issued' not in c]
print((df[stall_cols].sum() / tot_s * 100).sort_values(ascending=False).head(12).to_string())
# hot ranges: split code at big changes of execution count
W = 200
df['blk'] = df.index // W
b = df.groupby('blk').agg(inst=('Instructions Executed', 'sum'), samp=('# Samples', 'sum'))
b['inst%'] = 100 * b.inst / tot_inst; b['samp%'] = 100 * b.samp / tot_s
print(b[b['samp%'] > 1.0].to_string())
