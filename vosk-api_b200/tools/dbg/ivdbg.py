import sys, os, tempfile
sys.path[:0]=['/root/repo/vosk-api_b200','/root/repo/vosk-api_b200/tools','/root/repo/oracle','/root/repo/tests']
import numpy as np, vbmodel, oracle, helpers, vosk
vosk.SetLogLevel(-1)
td=tempfile.mkdtemp()
mdir=vbmodel.write_model_dir(td,"small",0)
model=vbmodel.load_model_dir(mdir)
waves=[vbmodel.synth_audio(s, 900+i) for i,s in enumerate([2.5,0.9])]
for opts in ("num-channels=4,max-batch-size=4,max-seconds=10","lattice=1,num-channels=4,max-batch-size=4,max-seconds=10"):
    got,_=helpers.run_engine(mdir,waves,options=opts)
    for w,g in zip(waves,got):
        ref=oracle.recognize(model,w,stages=True)
        iv=g["ivectors"].reshape(-1,40)
        print(opts[:9], iv.shape, "iv err per chunk", np.abs(iv-ref["ivectors"]).max(axis=1), "mfcc", np.abs(g["mfcc"]-ref["mfcc"]).max())
