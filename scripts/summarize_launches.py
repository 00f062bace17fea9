import csv, sys, collections
def main(path, last=None):
    rows=[l for l in open(path) if l.startswith('"')]
    r=list(csv.reader(rows)); hdr=r[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    seq=[(x[ki].split('(')[0][-50:], float(x[vi].replace(',',''))/1e6) for x in r[1:]]
    if last: seq=seq[-last:]
    for n,v in seq: print(f"{v:10.3f} ms  {n}")
main(sys.argv[1], int(sys.argv[2]) if len(sys.argv)>2 else None)
