import os, subprocess, sys
for so in sys.argv[1:]:
    env = dict(os.environ, PIXO_B200_SO=os.path.abspath(so))
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "huff_time.py")], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-400:])
