import os, subprocess, sys
for so in sys.argv[1:]:
    env = dict(os.environ, PIXO_B200_SO=os.path.abspath(so))
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "quick_png.py")], env=env, capture_output=True, text=True).stdout
    print(os.path.basename(so), "|", " | ".join(l.split(":")[1].strip().split("  ")[0] for l in out.strip().splitlines()[:6]))
