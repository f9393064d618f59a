"""Run the reference's own test files against the B200 method modules (last bullet).

One-off validation, not part of tests/ or bench.py: it needs the unmodified reference installed in the git-ignored
baseline/_ref (python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of
/root/reference>), which travels to the GPU box with the snapshot.  The selected test files are copied to a
scratch directory with the method names swapped ('PCGPwM' -> 'PCGPwM_B200', 'PCGPwImpute' -> 'PCGPwImpute_B200',
'directbayeswoodbury' -> 'directbayeswoodbury_B200', 'mlbayeswoodbury' -> 'mlbayeswoodbury_B200', 'indGP' -> 'indGP_B200', 'PCSK' -> 'PCSK_B200') and a conftest that calls surmise_b200.register(); pytest then runs them through
surmise's unmodified emulator / calibrator facades.

    python scripts/run_reference_suite.py [--orig]      # --orig: run the same files unmodified (CPU reference)
"""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
FILES = ['test_emu_pcgpwm.py', 'test_cal_directbayeswoodbury.py', 'test_new_cal.py', 'test_new_emu.py',
         'test_new_emu_prediction.py', 'test_emu_supplement.py', 'test_emu_predict.py', 'test_emu_update.py',
         'test_emu_remove.py', 'test_emu_param_str.py', 'test_emu_pcgpwimpute.py']


def main():
    orig = '--orig' in sys.argv
    src = os.path.join(REF, 'surmise', 'tests')
    if not os.path.isdir(src):
        print('reference not installed in baseline/_ref')
        return 2
    out = os.path.join(ROOT, 'gpurun_out', 'ref_suite_orig' if orig else 'ref_suite')
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    for name in FILES:
        text = open(os.path.join(src, name)).read()
        if not orig:
            text = re.sub(r"(['\"])PCGPwM\1", r"\1PCGPwM_B200\1", text)
            text = re.sub(r"(['\"])directbayeswoodbury\1", r"\1directbayeswoodbury_B200\1", text)
            text = re.sub(r"(['\"])PCGPwImpute\1", r"\1PCGPwImpute_B200\1", text)
            text = re.sub(r"(['\"])mlbayeswoodbury\1", r"\1mlbayeswoodbury_B200\1", text)
            text = re.sub(r"(['\"])indGP\1", r"\1indGP_B200\1", text)
            text = re.sub(r"(['\"])PCSK\1", r"\1PCSK_B200\1", text)
        open(os.path.join(out, name), 'w').write(text)
    with open(os.path.join(out, 'conftest.py'), 'w') as fh:
        fh.write('import sys\nsys.path.insert(0, %r)\nsys.path.insert(0, %r)\n' % (REF, ROOT))
        if not orig:
            fh.write('import surmise_b200\nsurmise_b200.register()\n')
    log = os.path.join(ROOT, 'gpurun_out', 'ref_suite_orig.txt' if orig else 'ref_suite.txt')
    cmd = [sys.executable, '-m', 'pytest', out, '-q', '-p', 'no:cacheprovider', '--timeout', '1200', '-rfE',
           '--tb=short', '-W', 'ignore']
    with open(log, 'w') as fh:
        rc = subprocess.call(cmd, stdout=fh, stderr=subprocess.STDOUT, cwd=out)
    tail = open(log).read().strip().splitlines()[-25:]
    print('\n'.join(tail))
    return rc


if __name__ == '__main__':
    sys.exit(main())
