"""Generates the golden fixtures in this directory from the reference checkout (/root/reference).

The reference's own tests hold no trajectory values; what it does pin for this path:
  * model sizes + masses            tests/test_flybare.py:12-36
  * walk_imitation obs names/action dim/timesteps   tests/test_walking_env.py:11-24,46,56-57
  * action order, ctrl ranges, obs shapes            docs/getting-started.ipynb (cells 44, 46)
Run:  python tests/golden/make_goldens.py
"""
import ast
import json
import os
import re

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    g = {}
    src = open(os.path.join(REF, 'tests/test_flybare.py')).read()
    g['flybare_sizes'] = ast.literal_eval(re.search(r'expect = (\{.*?\})', src, re.S).group(1))
    g['flybare_masses'] = ast.literal_eval(re.search(r'expect_close_masses = (\{.*?\})', src, re.S).group(1))
    src = open(os.path.join(REF, 'tests/test_walking_env.py')).read()
    names = ast.literal_eval(re.search(r'expect_obs_names = (\[.*?\])', src, re.S).group(1))
    g['walk_obs_names'] = ['walker/' + s for s in names]
    g['walk_num_act'] = int(re.search(r'expect_num_act = (\d+)', src).group(1))
    nb = json.load(open(os.path.join(REF, 'docs/getting-started.ipynb')))
    for c in nb['cells']:
        s = ''.join(c['source'])
        for o in c.get('outputs', []):
            txt = ''.join(o.get('data', {}).get('text/plain', []))
            if s.strip() == 'env.action_spec()':
                g['action_names'] = re.search(r"name='(.*?)'", txt, re.S).group(1).split('\\t')
                g['action_minimum'] = [float(x) for x in re.search(r'minimum=\[(.*?)\]', txt, re.S).group(1).split()]
                g['action_maximum'] = [float(x) for x in re.search(r'maximum=\[(.*?)\]', txt, re.S).group(1).split()]
            if s.strip() == 'env.observation_spec()':
                g['walk_on_ball_obs_shapes'] = {m.group(1): int(m.group(2)) for m in
                                                re.finditer(r"\('(walker/\w+)',\s*Array\(shape=\((\d+),\)", txt)}
    json.dump(g, open(os.path.join(OUT, 'reference_goldens.json'), 'w'), indent=1)
    print({k: (len(v) if hasattr(v, '__len__') else v) for k, v in g.items()})


if __name__ == '__main__':
    main()
