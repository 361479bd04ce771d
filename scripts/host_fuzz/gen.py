import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hypothesis import given, settings, seed, strategies as st
from tests.test_obj_loader import obj_text
from tests.test_scene_yaml import scene_text
os.makedirs("corpus", exist_ok=True)
rnd = random.Random(1)
n = [0]
def mutate(b):
    b = bytearray(b)
    for _ in range(rnd.randint(0, 6)):
        if not b: break
        k = rnd.randrange(len(b)); op = rnd.randrange(4)
        if op == 0: b[k] = rnd.choice(b" \t\n/-+.0123456789eE:{}[],'\"#&*~xvfl\x00\xff")
        elif op == 1: del b[k:k + rnd.randint(1, 8)]
        elif op == 2: b[k:k] = bytes(rnd.choice(b" \n/-:{}[],'\"#&*0123456789") for _ in range(rnd.randint(1, 6)))
        else: b[k:k] = b[max(0, k - 20):k]
    return bytes(b)
def emit(kind, text):
    raw = text.encode()
    for v in range(4):
        open(f"corpus/{kind}_{n[0]:05d}.{kind}", "wb").write(raw if v == 0 else mutate(raw)); n[0] += 1
@seed(11)
@settings(max_examples=400, deadline=None, database=None)
@given(t=obj_text())
def a(t): emit("obj", t)
@seed(12)
@settings(max_examples=400, deadline=None, database=None)
@given(t=scene_text())
def b(t): emit("yaml", t)
a(); b(); print(n[0])
