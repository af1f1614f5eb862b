"""Host tokenizer throughput: declip_b200.tokenizer (C++ BPE, threads) vs the reference's Python SimpleTokenizer (when
/root/reference is importable), on synthetic captions built from the toy corpus vocabulary.  python tools/bpe_perf.py"""
import json, os, random, sys, time, gzip, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from declip_b200.tokenizer import SimpleTokenizer

MERGES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bpe_small_merges.txt")
random.seed(0)
words = ("a photo of the big red dog running in park two dogs are playing with ball on grass quick brown fox jumps over lazy "
         "man riding bicycle down street woman holding umbrella rain it's beautiful day coffee tea 2023 children flower "
         "mountains lakes rivers sunset airplane flying blue sky people walking beach ocean waves photography").split()
caps = [" ".join(random.choice(words) for _ in range(random.randint(6, 24))) + "." for _ in range(16384)]
res = {}
for th in (1, 8, 16):
    tok = SimpleTokenizer(MERGES, threads=th)
    tok.tokenize(caps[:512], 77)
    t0 = time.perf_counter(); tok.tokenize(caps, 77); dt = time.perf_counter() - t0
    res["cpp_threads_%d" % th] = round(len(caps) / dt)
try:
    from oracle import ref_harness
    ref_harness.setup()
    from prototype.model.utils.text_utils.simple_tokenizer import SimpleTokenizer as Ref
    gz = os.path.join(tempfile.gettempdir(), "bpe_perf.txt.gz")
    with gzip.open(gz, "wt", encoding="utf-8") as f:
        f.write(open(MERGES, encoding="utf-8").read())
    ref = Ref(gz)
    sot, eot = ref.encoder["<|startoftext|>"], ref.encoder["<|endoftext|>"]
    [ref.encode(c) for c in caps[:512]]
    t0 = time.perf_counter(); [[sot] + ref.encode(c) + [eot] for c in caps[:4096]]; dt = time.perf_counter() - t0
    res["reference_python"] = round(4096 / dt)
except Exception as e:  # noqa: BLE001
    res["reference_python"] = "unavailable: %s" % e
print(json.dumps({"captions_per_second": res, "captions": len(caps), "host_cpus": os.cpu_count()}))
