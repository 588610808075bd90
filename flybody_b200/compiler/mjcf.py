"""MJCF-subset reader for the fruit-fly model (`fruitfly.xml`) + the reference's Python
model surgery, without dm_control.

What is covered is exactly the subset SURVEY.md App. C.1 enumerates for
`flybody/fruitfly/assets/fruitfly.xml`: nested default classes with `childclass`, bodies,
hinge/free joints, geoms (mesh/capsule/ellipsoid/cylinder/sphere/box/plane, `fromto`,
`euler`), sites, contact excludes, fixed tendons, `general`/`adhesion` actuators and
site sensors.

The edits replayed here on the XML tree are the ones the reference performs through PyMJCF:
  * `FruitFly._build`                         flybody/fruitfly/fruitfly.py:186-340
  * `FruitFlyTask.__init__` (attach, ghost)   flybody/tasks/base.py:130-167
  * `Walking.__init__` / `Flying.__init__`    flybody/tasks/base.py:296-364,385-428
  * `WalkImitation.__init__` (claw friction)  flybody/tasks/walk_imitation.py:70-73
"""
import copy
import xml.etree.ElementTree as ET

import numpy as np

from .quat import (axisangle2q, euler2q, neg_quat, qmul, qnorm, qrot, z2quat)

ACT_TAGS = ('general', 'adhesion', 'motor', 'position', 'velocity')
LEG_SUBSTR = ('T1', 'T2', 'T3')
MOUTH_SUBSTR = ('rostrum', 'haustellum', 'labrum')
LEG_BODY_SUBSTR = ('coxa', 'femur', 'tibia', 'tarsus', 'claw')


def _any_in(subs, s):
    return any(x in s for x in subs)


def fvec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    if isinstance(s, str):
        v = np.array([float(t) for t in s.split()], dtype=np.float64)
    else:
        v = np.atleast_1d(np.asarray(s, dtype=np.float64))
    if n is not None and len(v) < n:
        d = np.zeros(n) if default is None else np.array(default, dtype=np.float64)
        d[:len(v)] = v
        v = d
    return v


def sset(el, key, val):
    """Set an attribute from a number / sequence."""
    if isinstance(val, str):
        el.set(key, val)
    else:
        el.set(key, ' '.join(repr(float(x)) for x in np.atleast_1d(val)))


# ------------------------------------------------------------------------------------------
# XML tree helpers (PyMJCF-like finds)
# ------------------------------------------------------------------------------------------

class FlyXML:
    """Mutable MJCF tree of one fly with PyMJCF-style lookup helpers."""

    def __init__(self, xml_path):
        self.tree = ET.parse(xml_path)
        self.root = self.tree.getroot()
        self._reindex()

    def _reindex(self):
        self.parent = {c: p for p in self.root.iter() for c in p}

    @property
    def worldbody(self):
        return self.root.find('worldbody')

    def all(self, tag):
        if tag == 'actuator':
            sec = self.root.find('actuator')
            return list(sec) if sec is not None else []
        if tag == 'tendon':
            sec = self.root.find('tendon')
            return list(sec) if sec is not None else []
        if tag == 'sensor':
            sec = self.root.find('sensor')
            return list(sec) if sec is not None else []
        if tag == 'exclude':
            sec = self.root.find('contact')
            return list(sec) if sec is not None else []
        if tag == 'joint':
            return [e for e in self.worldbody.iter() if e.tag in ('joint', 'freejoint')]
        return [e for e in self.worldbody.iter() if e.tag == tag]

    def find(self, tag, name):
        for e in self.all(tag):
            if e.get('name') == name:
                return e
        return None

    def remove(self, el):
        self.parent[el].remove(el)
        self._reindex()

    def default_class(self, name):
        for d in self.root.find('default').iter('default'):
            if d.get('class') == name:
                return d
        if name == 'main':
            return self.root.find('default')
        return None

    def class_child(self, cls_name, tag, create=True):
        d = self.default_class(cls_name)
        for ch in d:
            if ch.tag == tag:
                return ch
        if create:
            return ET.SubElement(d, tag)
        return None

    def class_parent(self, cls_name):
        d = self.default_class(cls_name)
        p = self.parent.get(d)
        return p

    # own (non-inherited) attribute of a joint's explicit class, PyMJCF `joint.dclass.joint.<attr>`
    def own_class_attr(self, cls_name, tag, attr):
        d = self.default_class(cls_name)
        if d is None:
            return None
        for ch in d:
            if ch.tag == tag and ch.get(attr) is not None:
                return ch.get(attr)
        return None


def body_quat_from_springrefs(fx, body):
    """reference `fruitfly.py:68-87`."""
    joints = [j for j in body if j.tag == 'joint']
    if not joints:
        return None
    quats = []
    for j in joints:
        cls = j.get('class')
        theta = j.get('springref') or (fx.own_class_attr(cls, 'joint', 'springref') if cls else None) or 0
        theta = float(theta)
        axis = j.get('axis') or (fx.own_class_attr(cls, 'joint', 'axis') if cls else None)
        if axis is None:
            pcls = fx.class_parent(cls).get('class')
            axis = fx.own_class_attr(pcls, 'joint', 'axis')
        axis = fvec(axis)
        quats.append(np.hstack((np.cos(theta / 2), np.sin(theta / 2) * axis)))
    quat = np.array([1.0, 0, 0, 0])
    for i in range(len(quats)):
        quat = qmul(quats[-1 - i], quat)
    if body.get('quat') is not None:
        quat = qmul(fvec(body.get('quat')), quat)
    return quat


def change_body_frame(fx, body, frame_pos, frame_quat):
    """reference `fruitfly.py:90-114` (children keep their world placement; joint axes do not
    get rotated because joints have no quat attribute)."""
    frame_pos = np.zeros(3) if frame_pos is None else np.asarray(frame_pos, dtype=np.float64)
    frame_quat = np.array((1.0, 0, 0, 0)) if frame_quat is None else frame_quat
    body_pos = fvec(body.get('pos'), 3, (0, 0, 0))
    dpos = body_pos - frame_pos
    body_quat = fvec(body.get('quat'), 4, (1, 0, 0, 0))
    dquat = qmul(neg_quat(frame_quat), body_quat)
    sset(body, 'pos', frame_pos)
    sset(body, 'quat', frame_quat)
    for child in list(body):
        if child.tag not in ('geom', 'site', 'body', 'joint', 'camera', 'light', 'inertial'):
            continue
        if child.tag in ('geom', 'site', 'body', 'camera', 'inertial'):
            cq = fvec(child.get('quat'), 4, (1, 0, 0, 0))
            sset(child, 'quat', qmul(dquat, cq))
        cp = fvec(child.get('pos'), 3, (0, 0, 0))
        # mju_rotVecQuat / mju_mulQuat do not normalise: reproduce with un-normalised algebra
        pos_in_parent = _rot_unnorm(cp, body_quat) + dpos
        sset(child, 'pos', _rot_unnorm(pos_in_parent, neg_quat(frame_quat)))


def _rot_unnorm(v, q):
    """mju_rotVecQuat: v' = q v q^-1 computed as matrix of (possibly un-normalised) q."""
    from .quat import q2mat
    return q2mat(q) @ v


def build_fly_xml(xml_path, name='walker', use_legs=True, use_wings=False, use_mouth=False,
                  use_antennae=False, force_actuators=False, joint_filter=0.01,
                  adhesion_filter=0.007, dyntype_filterexact=False, body_pitch_angle=47.5,
                  stroke_plane_angle=0.0):
    """Replays `FruitFly._build` (reference `fruitfly.py:181-340`) on the XML tree."""
    fx = FlyXML(xml_path)
    root = fx.root
    fx.name = name
    # Remove freejoint (fruitfly.py:187); the task re-adds a free joint on the attachment frame.
    fj = fx.find('joint', 'free')
    if fj is not None:
        fx.remove(fj)

    observable_joints = [j.get('name') for j in fx.all('joint')]

    def act_by_name(n):
        return fx.find('actuator', n)

    if not use_legs:
        for body in fx.all('body'):
            if _any_in(LEG_SUBSTR, body.get('name')):
                q = body_quat_from_springrefs(fx, body)
                if q is not None:
                    sset(body, 'quat', q)
        for tendon in fx.all('tendon'):
            if _any_in(LEG_SUBSTR, tendon.get('name')):
                a = act_by_name(tendon.get('name'))
                if a is not None:
                    fx.remove(a)
                fx.remove(tendon)
        for joint in fx.all('joint'):
            if _any_in(LEG_SUBSTR, joint.get('name')):
                a = act_by_name(joint.get('name'))
                if a is not None:
                    fx.remove(a)
                observable_joints.remove(joint.get('name'))
                fx.remove(joint)
        for a in fx.all('actuator'):
            if 'adhere' in a.get('name') and _any_in(LEG_SUBSTR, a.get('name')):
                fx.remove(a)
        for s in fx.all('sensor'):
            if _any_in(LEG_SUBSTR, s.get('name')):
                fx.remove(s)

    if not use_wings:
        for joint in fx.all('joint'):
            if 'wing' in joint.get('name'):
                fx.remove(act_by_name(joint.get('name')))
                observable_joints.remove(joint.get('name'))
        for s in fx.all('sensor'):
            if 'wing' in s.get('name'):
                fx.remove(s)

    if not use_mouth:
        for joint in fx.all('joint'):
            if _any_in(MOUTH_SUBSTR, joint.get('name')):
                fx.remove(act_by_name(joint.get('name')))
                observable_joints.remove(joint.get('name'))
        for a in fx.all('actuator'):
            if 'adhere' in a.get('name') and _any_in(MOUTH_SUBSTR, a.get('name')):
                fx.remove(a)

    if not use_antennae:
        for joint in fx.all('joint'):
            if 'antenna' in joint.get('name'):
                fx.remove(act_by_name(joint.get('name')))
                observable_joints.remove(joint.get('name'))

    if use_wings:
        up_site = fx.find('site', 'hover_up_dir')
        up_dir = fvec(up_site.get('quat'))
        up_dir_angle = 2 * np.arccos(up_dir[0])
        delta = np.deg2rad(body_pitch_angle) - up_dir_angle
        dquat = np.array([np.cos(delta / 2), 0, np.sin(delta / 2), 0])
        up_dir = qmul(dquat, up_dir)
        sset(up_site, 'quat', up_dir)
        spa = np.deg2rad(stroke_plane_angle)
        sp_quat = np.array([np.cos(spa / 2), 0, np.sin(spa / 2), 0])
        for quat, wing in [(np.array([0.0, 0, 0, 1]), 'wing_left'),
                           (np.array([0.0, -1, 0, 0]), 'wing_right')]:
            dq = qmul(neg_quat(sp_quat), quat)
            new_wing_quat = qmul(dq, neg_quat(up_dir))
            body = fx.find('body', wing)
            change_body_frame(fx, body, fvec(body.get('pos'), 3, (0, 0, 0)), new_wing_quat)

    if force_actuators:
        for d in root.find('default').iter('default'):
            for ch in d:
                if ch.tag == 'general':
                    for k in ('biastype', 'biasprm', 'ctrlrange'):
                        ch.attrib.pop(k, None)
        fx.class_child('main', 'general').set('ctrlrange', '-1 1')
        for a in fx.all('actuator'):
            if a.tag == 'adhesion':
                continue
            for k in ('biastype', 'biasprm', 'ctrlrange'):
                a.attrib.pop(k, None)

    dyntype = 'filterexact' if dyntype_filterexact else 'filter'
    if joint_filter > 0:
        for a in fx.all('actuator'):
            if a.tag != 'adhesion':
                a.set('dyntype', dyntype)
                a.set('dynprm', repr(float(joint_filter)))
    if adhesion_filter > 0:
        for a in fx.all('actuator'):
            if a.tag == 'adhesion':
                pcls = fx.class_parent(a.get('class')).get('class')
                g = fx.class_child(pcls, 'general')
                g.set('dyntype', dyntype)
                g.set('dynprm', repr(float(adhesion_filter)))

    # action-class <-> ctrl index maps (fruitfly.py:342-379)
    name_substr = {'adhesion': [], 'head': ['head'], 'mouth': list(MOUTH_SUBSTR),
                   'antennae': ['antenna'], 'wings': ['wing'], 'abdomen': ['abdomen'],
                   'legs': list(LEG_SUBSTR), 'user': []}
    names = [a.get('name') for a in fx.all('actuator')]
    ctrl_indices = {}
    for cls, subs in name_substr.items():
        idx = [i for i, n in enumerate(names) if _any_in(subs, n) and 'adhere' not in n]
        ctrl_indices[cls] = idx if idx else None
    idx = [i for i, n in enumerate(names) if 'adhere' in n]
    ctrl_indices['adhesion'] = idx if idx else None
    fx.ctrl_indices = ctrl_indices
    fx.observable_joints = observable_joints
    return fx


# ------------------------------------------------------------------------------------------
# Default-class resolution
# ------------------------------------------------------------------------------------------

def parse_defaults(root):
    classes = {}

    def rec(node, parent_cls):
        name = node.get('class', 'main')
        cls = copy.deepcopy(parent_cls) if parent_cls is not None else {}
        for ch in node:
            if ch.tag == 'default':
                continue
            tag = 'actuator' if ch.tag in ACT_TAGS else ch.tag
            attrs = dict(ch.attrib)
            if ch.tag == 'adhesion' and 'gain' in attrs:
                attrs['gainprm'] = attrs.pop('gain')
            cls.setdefault(tag, {}).update(attrs)
        classes[name] = cls
        for ch in node:
            if ch.tag == 'default':
                rec(ch, cls)

    d = root.find('default')
    if d is not None:
        rec(d, None)
    else:
        classes['main'] = {}
    return classes


def resolved(el, classes, childclass, tag=None):
    """Attributes of `el` with defaults applied (explicit class > childclass > main)."""
    tag = tag or ('actuator' if el.tag in ACT_TAGS else el.tag)
    cls = el.get('class') or childclass or 'main'
    out = dict(classes.get(cls, {}).get(tag, {}))
    attrs = dict(el.attrib)
    if el.tag == 'adhesion' and 'gain' in attrs:
        attrs['gainprm'] = attrs.pop('gain')
    out.update(attrs)
    return out


def geom_frame(a):
    """pos, quat, size of a geom/site from resolved attributes (handles fromto / euler)."""
    size = fvec(a.get('size'), 3, (0, 0, 0)) if a.get('size') is not None else np.zeros(3)
    if a.get('fromto') is not None:
        ft = fvec(a['fromto'])
        p0, p1 = ft[:3], ft[3:]
        vec = p0 - p1                      # MuJoCo: z axis points from "to" to "from"
        pos = 0.5 * (p0 + p1)
        quat = z2quat(vec)
        size = np.array([size[0], 0.5 * np.linalg.norm(vec), 0.0])
        return pos, quat, size
    pos = fvec(a.get('pos'), 3, (0, 0, 0))
    if a.get('euler') is not None:
        quat = euler2q(fvec(a['euler']))
    else:
        quat = qnorm(fvec(a.get('quat'), 4, (1, 0, 0, 0)))
    return pos, quat, size
