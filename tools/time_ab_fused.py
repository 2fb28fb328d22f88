"""In-solver generator (fused mode 1) of several builds of the library: tools/time_ab_fused.py lib1.so ..."""
import os, subprocess, sys
for lib in sys.argv[1:]:
    env = dict(os.environ, DDB_FUSED_INKERNEL='1')
    if lib != 'default':
        env['DDB200_LIBRARY'] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, 'tools/fused_ab.py'], env=env, capture_output=True, text=True)
    print(lib, r.stdout.strip()[-90:], r.stderr.strip()[-200:], flush=True)
