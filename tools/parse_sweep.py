import csv, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
sizes=[50_000,200_000,709_000,2_900_000,12_000_000,48_000_000]
i=0
for n in sizes:
    blk=rows[i:i+16]; i+=16
    d=[float(r['Metric Value'].replace(',',''))/1000 for r in blk]
    u32=d[5:10]; u16=d[13:16]
    b32=n*(4+4*16)/1e3/sum(u32) ; b16=n*(2+2*12)/1e3/sum(u16)
    print(f"n={n:>9}: u32 hist {u32[0]:7.1f} passes {[round(x,1) for x in u32[1:]]} total {sum(u32):8.1f} us -> {b32:7.1f} GB/s | u16 hist {u16[0]:6.1f} passes {[round(x,1) for x in u16[1:]]} total {sum(u16):7.1f} us -> {b16:7.1f} GB/s")
