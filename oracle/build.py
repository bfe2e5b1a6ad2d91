"""Compile oracle/c/oracle.c -> oracle/_build/liboracle.so (gcc, IEEE-strict flags).

TEST INFRASTRUCTURE ONLY.  Called by __graft_entry__.build(); the .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'c', 'oracle.c')
OUT_DIR = os.path.join(HERE, '_build')
OUT = os.path.join(OUT_DIR, 'liboracle.so')


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ['gcc', '-O2', '-std=c99', '-fPIC', '-shared', '-ffp-contract=off', '-fno-fast-math',
           '-Wall', '-o', OUT, SRC, '-lm']
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force=True))
