import json,sys,os
sys.path.insert(0,os.getcwd())
import torch, rfanalyzer_b200 as rfa
S=1<<24
stream=torch.cuda.Stream(); ctx=rfa.Context(0,stream)
for variant in (0,2):
    ctx.set_option("rs_span", variant)
    for name,fmt,fs,mode,width,packet in (("nfm",2,10_000_000,2,10_000,65536),("cw",2,10_000_000,6,300,65536),("wfm20M",0,20_000_000,3,100_000,131072)):
        off=fs//10
        with torch.cuda.stream(stream):
            iq=torch.empty(S*rfa.BYTES_PER_SAMPLE[fmt],dtype=torch.uint8,device="cuda"); rfa.synth_iq(ctx,fmt,S,iq)
            plan=rfa.ChainPlan(ctx,fmt,fs,100_000_000,100_000_000+off,mode,width,packet,1.0,rfa.SUM_FMA)
            audio=torch.empty(plan.max_audio(S),dtype=torch.float32,device="cuda")
            plan.process(iq,S,audio); stream.synchronize()
            a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(20): plan.process(iq,S,audio)
            b.record(stream); stream.synchronize()
            print("variant",variant,name,"I/D",plan.interpolation,plan.decimation,round(a.elapsed_time(b)/20*1e3,1),"us")
