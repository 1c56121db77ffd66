import json, os, sys, torch
sys.path.insert(0, os.getcwd())
sys.argv=[sys.argv[0]]
import scripts.bench_configs as bc
for K,T,g in ((32768,50,256),(65536,50,256),(131072,50,512),(16384,50,256)):
    r=bc.single_case(K,T,g); print(json.dumps({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items()}))
r=bc.stoch_case(); print(json.dumps({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items()}))
