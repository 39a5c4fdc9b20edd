import sys, os
sys.path.insert(0, "/root/repo")
sys.argv=[sys.argv[0]]
from tools import gpu_check as G
from object_detection_destr_b200 import _lib
def knob(i,v): _lib.lib.destr_debug_knob(i,v)
print("== default"); G.attn_case(8,1050,8,"random",True)
print("== B=1 (72 items, 1 per CTA)"); G.attn_case(1,1050,8,"random",False)
print("== B=4 (288 items, 1 per CTA, 2 CTAs/SM)"); G.attn_case(4,1050,8,"random",False)
print("== B=5 (360 items)"); G.attn_case(5,1050,8,"random",False)
print("== tau=0 (always rescale)"); knob(11,1); G.attn_case(8,1050,8,"random",True)
print("== tau=1000 (never)"); knob(11,1001); G.attn_case(8,1050,8,"random",True)
knob(11,0)
print("== grid=148"); knob(12,148); G.attn_case(8,1050,8,"random",True)
print("== grid=64"); knob(12,64); G.attn_case(8,1050,8,"random",True)
