"""yaml config -> EnvSpec compiler (host side only).

Keeps the reference's yaml schema (`General / Agents / Entities / Rules`,
marl_factory_grid/utils/config_parser.py:16-274) but instead of importing Python classes by name it
resolves every built-in name to an opcode / table entry that the CUDA kernels interpret.
Names that the reference would look up in a user `custom_modules_path` cannot run on the GPU and
raise `NotImplementedError`.  Where the reference prints and calls `exit(-99999)`
(config_parser.py:121, 245) this module raises `ValueError` instead.

Several yaml keys are accepted and deliberately ignored because the reference swallows them too
(defect B8): e.g. `Batteries.initial_charge`, `DirtPiles.dirt_spawn_r_var`,
`DropOffLocations.max_dropoff_storage_size`, `Destinations.spawn_mode`, `DoorAutoClose.close_frequency`.
"""
from __future__ import annotations

import ast
from pathlib import Path
from typing import Dict, List, Union

import numpy as np
import yaml

from . import spec as S
from .level_parser import LevelParser, resolve_level_path
from .spec import ActionSpec, AgentSpec, ChannelSpec, EnvSpec, GroupSpec, RuleSpec

# reward defaults: environment/rewards.py:1-5, modules/*/constants.py
_MOVE_VALID, _MOVE_FAIL, _NOOP = -0.001, -0.05, -0.01

# name in yaml -> (Action.name, class name, opcode, default valid reward, default fail reward)
_MODULE_ACTIONS = {
    'Noop': ('Noop', 'Noop', S.OP_NOOP, _NOOP, _NOOP),
    'DoorUse': ('use_door', 'DoorUse', S.OP_DOORUSE, -0.0, -0.01),                # doors/constants.py:21-22
    'Clean': ('do_cleanup_action', 'Clean', S.OP_CLEAN, 0.5, -0.1),               # clean_up/constants.py:9-10
    # valid/fail defaults are swapped in the reference (items/actions.py:21): success -> -0.1, failure -> +0.1
    'ItemAction': ('ITEMACTION', 'ItemAction', S.OP_ITEM, -0.1, 0.1),
    'Charge': ('do_charge_action', 'Charge', S.OP_CHARGE, 0.1, -0.1),             # batteries/constants.py:9-10
    'DestAction': ('Destinations', 'DestAction', S.OP_DEST, 0.1, -0.1),           # destinations/constants.py:12-13
    'MachineAction': ('Maintain', 'MachineAction', S.OP_MACHINE, 0.5, -0.1),      # machines/constants.py
}
_ACTIONS_WITHOUT_KWARGS = {'Clean', 'Charge', 'DestAction', 'MachineAction'}       # their __init__ takes no kwargs

_PER_AGENT_GROUPS = ('Batteries', 'Inventories', 'GlobalPositions')
_QUANTITY_GROUPS = ('ChargePods', 'Destinations', 'Items', 'DropOffLocations', 'Machines', 'Maintainers')


def _n_abbr(n: int) -> str:
    return {1: 'st', 2: 'nd', 3: 'rd'}.get(n, 'th')            # config_parser.py:41-45


class FactoryConfigParser:
    default_actions = ['Move8', 'Noop']                          # config_parser.py:19
    default_observations = ['Walls', 'Agent']                    # config_parser.py:20

    def __init__(self, config_path: Union[str, Path], custom_modules_path=None):
        if custom_modules_path is not None:
            raise NotImplementedError('custom_modules_path: user-defined Python modules cannot be compiled to the '
                                      'B200 engine; only the built-in actions / rules / entities are supported.')
        self.config_path = Path(config_path)
        with self.config_path.open() as fh:
            self.config = yaml.safe_load(fh)
        for section in ('General', 'Agents', 'Entities', 'Rules'):
            if section not in self.config or self.config[section] is None:
                if section == 'General' or section == 'Agents':
                    raise ValueError(f'The mandatory "{section}" section could not be found in {self.config_path}.')
                self.config[section] = {}

    # ---- reference-compatible accessors (config_parser.py:34-78)
    def __getattr__(self, item):
        try:
            return self.__dict__['config']['General'][item]
        except KeyError:
            raise AttributeError(item)

    def __getitem__(self, item):
        return self.config[item]

    @property
    def agents(self):
        return self.config['Agents']

    @property
    def entities(self):
        return self.config['Entities']

    @property
    def rules(self):
        return self.config['Rules']

    # ------------------------------------------------------------------------------------ actions
    def _parse_actions(self, agent_name: str, conf_actions) -> List[ActionSpec]:
        if isinstance(conf_actions, dict):
            conf_kwargs = {k: (v or {}) for k, v in conf_actions.items()}
            conf_actions = list(conf_actions.keys())
        elif isinstance(conf_actions, list):
            conf_kwargs = {}
            if any(isinstance(x, dict) for x in conf_actions):
                raise ValueError(f'Agent {agent_name}: per-action kwargs need the dict form of "Actions".')
        else:
            raise ValueError(f'Agent {agent_name}: "Actions" must be a list or a dict.')
        names = []
        for action in conf_actions:
            names.extend(self.default_actions if action == 'Defaults' else [action])
        out: List[ActionSpec] = []
        for action in names:
            kw = dict(conf_kwargs.get(action, {}))
            if action in ('Move8', 'Move4'):
                dirs = range(8) if action == 'Move8' else range(4)
                valid, fail = _pick(kw, 'valid_reward', _MOVE_VALID), _pick(kw, 'fail_reward', _MOVE_FAIL)
                for d in dirs:
                    out.append(ActionSpec(S.DIR_NAMES[d], S.DIR_CLASS[d], S.OP_MOVE, d, valid, fail))
                _no_extra(kw, ('valid_reward', 'fail_reward'), action)
            elif action in _MODULE_ACTIONS:
                name, cls, op, dv, df = _MODULE_ACTIONS[action]
                if kw and action in _ACTIONS_WITHOUT_KWARGS:
                    raise TypeError(f'{action}.__init__() takes no keyword arguments (got {sorted(kw)}).')
                aux = 0.0
                allowed = ['valid_reward', 'fail_reward']
                if action == 'ItemAction':
                    aux = _pick(kw, 'failed_dropoff_reward', -0.1)          # items/constants.py:12
                    allowed += ['failed_dropoff_reward', 'valid_dropoff_reward']
                out.append(ActionSpec(name, cls, op, 0, _pick(kw, 'valid_reward', dv), _pick(kw, 'fail_reward', df), aux))
                _no_extra(kw, allowed, action)
            else:
                raise NotImplementedError(f'Action "{action}" is not a built-in action of marl-factory-grid; custom '
                                          f'actions cannot be compiled to the B200 engine.')
        if len(out) > S.MAX_ACTIONS:
            raise ValueError(f'Agent {agent_name}: {len(out)} actions exceed the supported maximum {S.MAX_ACTIONS}.')
        return out

    # ------------------------------------------------------------------------------------ agents
    def parse_agents_conf(self) -> Dict[str, dict]:
        """Same expansion as config_parser.py:128-199 (Defaults first, Clones naming)."""
        parsed = {}
        for name, conf in self.agents.items():
            conf = conf or {}
            if conf.get('Observations') is None:
                raise AssertionError('Did you specify any Observation?')
            observations = []
            if 'Defaults' in conf['Observations']:
                observations.extend(self.default_observations)
            observations.extend(x for x in conf['Observations'] if x != 'Defaults')
            positions = [tuple(ast.literal_eval(x)) if isinstance(x, str) else tuple(x)
                         for x in (conf.get('Positions') or [])]
            other = {k: v for k, v in conf.items() if k not in ('Actions', 'Observations', 'Positions', 'Clones')}
            unknown = set(other) - {'is_blocking_pos'}
            if unknown:
                raise TypeError(f'Agent {name}: unsupported keyword(s) {sorted(unknown)}.')
            entry = dict(actions=self._parse_actions(name, conf['Actions']), observations=observations,
                         positions=positions, other=other)
            parsed[name] = entry
            clones = conf.get('Clones', 0)
            if clones:
                if isinstance(clones, int):
                    clones = [f'{name}_the_{n}{_n_abbr(n)}' for n in range(clones)]
                for clone in clones:
                    parsed[clone] = entry
        return parsed

    # ------------------------------------------------------------------------------------ compile
    def compile(self, custom_level_path=None, dirt_slots: int = 40) -> EnvSpec:
        general = self.config['General']
        pomdp_r = int(general.get('pomdp_r', 0) or 0)
        level_name = general.get('level_name', 'custom')
        level = LevelParser(resolve_level_path(level_name, custom_level_path))
        H, W = level.level_shape

        es = EnvSpec(level_name=level_name, H=int(H), W=int(W), walls=level.walls.copy(), floor=level.floor.copy(),
                     door_pos=np.zeros((0, 2), np.int32), pomdp_r=pomdp_r,
                     env_seed=int(general.get('env_seed', 69)),
                     individual_rewards=bool(general.get('individual_rewards', False)),
                     agents=[], rules=[], groups=[], dirt_slots=int(dirt_slots))

        self._pending_dest_rule = None
        # ---- Entities (yaml order = spawn order, config_parser.py:80-126 + level_parser.py:62-102)
        for gname, kwargs in (self.entities or {}).items():
            kwargs = dict(kwargs or {})
            if gname == 'Defaults':
                continue
            g = GroupSpec(gname)
            if gname == 'Doors':
                if len(level.door_pos) == 0:
                    raise ValueError("No Doors (Symbol: D) could be found!\nCheck your level file!")
                es.has_doors, es.door_pos = True, level.door_pos.copy()
                g.quantity = len(level.door_pos)
            elif gname == 'DirtPiles':
                es.has_dirt = True
                q = kwargs.get('coords_or_quantity', 10)
                if not isinstance(q, int):
                    raise NotImplementedError('DirtPiles.coords_or_quantity must be an integer quantity.')
                es.dirt_quantity = g.quantity = int(q)
                es.dirt_initial_amount = float(kwargs.get('initial_amount', 2))
                es.dirt_clean_amount = float(kwargs.get('clean_amount', 1))
                es.dirt_max_global = float(kwargs.get('max_global_amount', 20))
                es.dirt_n_var = float(kwargs.get('n_var', 0.2))
                es.dirt_amount_var = float(kwargs.get('amount_var', 0.2))
            elif gname in _PER_AGENT_GROUPS:
                if gname == 'Batteries':
                    es.has_batteries = True
                    es.battery_initial = float(kwargs.get('initial_charge_level', 1.0))
                elif gname == 'Inventories':
                    es.has_inventories = True
                else:
                    es.has_globalpos = True
            elif gname == 'Destinations' and kwargs.get('spawnrule'):
                # a custom spawn rule replaces the group's SpawnEntity rule (groups/collection.py:70-80)
                self._pending_dest_rule = dict(kwargs['spawnrule'])
                if len(self._pending_dest_rule) != 1 or next(iter(self._pending_dest_rule)) not in (
                        'SpawnDestinationOnAgent', 'SpawnDestinationsPerAgent'):
                    raise NotImplementedError(f'Destinations.spawnrule {sorted(self._pending_dest_rule)} is not supported.')
                es.n_dest = -1                                    # resolved once the agents are known
            elif gname in _QUANTITY_GROUPS:
                q = kwargs.get('coords_or_quantity', None)
                if isinstance(q, int) and not isinstance(q, bool):
                    g.quantity = int(q)
                elif isinstance(q, (list, tuple)) and q:
                    g.coords = [tuple(ast.literal_eval(x)) if isinstance(x, str) else tuple(x) for x in q]
                    g.quantity = len(g.coords)
                else:
                    raise ValueError(f'Entities.{gname} needs "coords_or_quantity" (int or list of coordinates).')
                if g.quantity > S.MAX_SMALL_GROUP:
                    raise ValueError(f'Entities.{gname}: {g.quantity} exceeds the supported maximum {S.MAX_SMALL_GROUP}.')
                attr = {'ChargePods': 'n_pods', 'Destinations': 'n_dest', 'Items': 'n_items',
                        'DropOffLocations': 'n_dropoff', 'Machines': 'n_machines', 'Maintainers': 'n_maint'}[gname]
                setattr(es, attr, g.quantity)
            else:
                raise NotImplementedError(f'Entity group "{gname}" is not a built-in group of marl-factory-grid; '
                                          f'custom entities cannot be compiled to the B200 engine.')
            es.groups.append(g)
        if es.n_doors > S.MAX_DOORS:
            raise ValueError(f'Level has {es.n_doors} doors; at most {S.MAX_DOORS} are supported.')
        if not 1 <= es.dirt_slots <= S.MAX_DIRT:
            raise ValueError(f'dirt_slots must be in [1, {S.MAX_DIRT}].')

        # ---- Agents
        parsed = self.parse_agents_conf()
        if not 1 <= len(parsed) <= S.MAX_AGENTS:
            raise ValueError(f'{len(parsed)} agents configured; supported: 1..{S.MAX_AGENTS}.')
        agent_names = [f'Agent[{n}]' for n in parsed]
        for idx, (name, conf) in enumerate(parsed.items()):
            channels = self._compile_channels(idx, agent_names, conf['observations'], es)
            if len(channels) > S.MAX_CHANNELS:
                raise ValueError(f'Agent {name}: {len(channels)} observation channels exceed {S.MAX_CHANNELS}.')
            for pos in conf['positions']:
                if not (0 <= pos[0] < H and 0 <= pos[1] < W) or level.walls[pos[0], pos[1]]:
                    raise ValueError(f'Agent {name}: position {pos} is not a floor tile of level {level_name}.')
            es.agents.append(AgentSpec(agent_names[idx], list(conf['actions']), channels, list(conf['positions']),
                                       bool(conf['other'].get('is_blocking_pos', False))))
        self._compile_dest_spawnrule(es, list(parsed))
        self._check_actions_vs_groups(es)

        # ---- Rules (yaml order; config_parser.py:201-250)
        for rname, kwargs in (self.rules or {}).items():
            es.rules.append(self._compile_rule(rname, dict(kwargs or {}), es))
        if len(es.rules) > S.MAX_RULES:
            raise ValueError(f'{len(es.rules)} rules exceed the supported maximum {S.MAX_RULES}.')
        return es

    # ------------------------------------------------------------------------------------ observations
    @staticmethod
    def _compile_channels(idx: int, agent_names: List[str], observations, es: EnvSpec) -> List[ChannelSpec]:
        """observation_builder.py:237-277 (layer naming) + :164-220 (how each name is resolved per step)."""
        me = agent_names[idx]
        others = [j for j in range(len(agent_names)) if j != idx]
        out: List[ChannelSpec] = []

        def group_term(name: str) -> int:
            if name not in S.GROUP_NAMES:
                raise ValueError(f'# No combination of "{name}" and "{me}" could be found in the observation sources.')
            present = {'Walls': True, 'Doors': es.has_doors, 'DirtPiles': es.has_dirt, 'Items': es.n_items > 0,
                       'DropOffLocations': es.n_dropoff > 0, 'ChargePods': es.n_pods > 0, 'Destinations': es.n_dest != 0,
                       'Machines': es.n_machines > 0, 'Maintainers': es.n_maint > 0}[name]
            if not present:
                raise ValueError(f'Observation "{name}" requested by {me} but "{name}" is not in Entities.')
            return S.GROUP_NAMES[name]

        for obs in observations:
            vals = None
            if isinstance(obs, dict):
                obs, vals = next(iter(obs.items()))
            if obs == 'Self':
                out.append(ChannelSpec(me, S.CH_TERMS, [S.G_AGENT0 + idx]))
            elif obs == 'Combined':
                vals = [vals] if isinstance(vals, str) else list(vals or [])
                terms = []
                for v in vals:
                    if v == 'Self':
                        terms.append(S.G_AGENT0 + idx)
                    elif v == 'Other':
                        terms.extend(S.G_AGENT0 + j for j in others)
                    elif v in ('Placeholder', 'Battery', 'Inventory', 'GlobalPosition'):
                        continue            # never a key of pre_sort_obs -> contributes nothing
                    else:
                        terms.append(group_term(v))
                out.append(ChannelSpec(f'Combined({me})', S.CH_TERMS, terms))
            elif obs == 'Other':
                out.extend(ChannelSpec(agent_names[j], S.CH_TERMS, [S.G_AGENT0 + j]) for j in others)
            elif obs == 'Agent':
                out.extend(ChannelSpec(agent_names[j], S.CH_TERMS, [S.G_AGENT0 + j]) for j in range(len(agent_names)))
            elif obs == 'Placeholder':
                out.append(ChannelSpec(obs, S.CH_ZERO))
            elif obs == 'Battery':
                if not es.has_batteries:
                    raise ValueError(f'Observation "Battery" requested by {me} but "Batteries" is not in Entities.')
                out.append(ChannelSpec(obs, S.CH_BATTERY))
            elif obs == 'GlobalPosition':
                if not es.has_globalpos:
                    raise ValueError(f'Observation "GlobalPosition" requested by {me} but "GlobalPositions" is missing.')
                out.append(ChannelSpec(obs, S.CH_GLOBALPOS))
            elif obs == 'Destination':
                # the singular name resolves (by the bound-entity regex, observation_builder.py:176-185) to the agent's own
                # bound Destination, which is positional => the plane is left untouched: always zeros [verified on the
                # reference's eight_puzzle / narrow_corridor scenarios]
                if es.n_dest == 0:
                    raise ValueError(f'# No combination of "Destination" and "{me}" could be found in the observation sources.')
                out.append(ChannelSpec(obs, S.CH_ZERO))
            elif obs == 'Inventory':
                if not es.has_inventories:
                    raise ValueError(f'Observation "Inventory" requested by {me} but "Inventories" is not in Entities.')
                out.append(ChannelSpec(obs, S.CH_ZERO))      # always zeros in the reference
            else:
                out.append(ChannelSpec(obs, S.CH_TERMS, [group_term(obs)]))
        return out

    def _compile_dest_spawnrule(self, es: EnvSpec, conf_names: List[str]):
        """SpawnDestinationOnAgent / SpawnDestinationsPerAgent (modules/destinations/rules.py:95-162): one destination per agent
        (resp. per dict entry), bound to that agent."""
        rule = self._pending_dest_rule
        es.dest_bound = [-1] * es.n_dest
        if not rule:
            return
        name, kw = next(iter(rule.items()))
        kw = dict(kw or {})
        if name == 'SpawnDestinationOnAgent':
            if kw:
                raise TypeError(f'SpawnDestinationOnAgent.__init__() got unexpected keyword argument(s) {sorted(kw)}.')
            es.dest_mode = S.DEST_ON_AGENT
            es.n_dest = es.n_agents
            es.dest_bound = list(range(es.n_agents))
            es.dest_cands = [[] for _ in range(es.n_agents)]
        else:
            per_agent = kw.pop('coords_or_quantity', None)
            if kw or not isinstance(per_agent, dict) or not per_agent:
                raise TypeError('SpawnDestinationsPerAgent needs coords_or_quantity: {agent name: [coordinates] | int}.')
            es.dest_mode = S.DEST_PER_AGENT
            es.dest_bound, es.dest_cands = [], []
            for agent_name, value in per_agent.items():
                # `h.get_first(state[c.AGENT], lambda x: agent_name in x.name)`: first agent whose name contains the key
                idx = next((i for i, a in enumerate(es.agents) if str(agent_name) in a.name), None)
                if idx is None:
                    raise AssertionError(f'SpawnDestinationsPerAgent: no agent matches "{agent_name}".')
                es.dest_bound.append(idx)
                if isinstance(value, int):
                    es.dest_cands.append([])                      # any floor tile (one destination, rules.py:127-143)
                else:
                    es.dest_cands.append([tuple(ast.literal_eval(x)) if isinstance(x, str) else tuple(x) for x in value])
            es.n_dest = len(es.dest_bound)
        if es.n_dest > S.MAX_SMALL_GROUP:
            raise ValueError(f'Entities.Destinations: {es.n_dest} exceeds the supported maximum {S.MAX_SMALL_GROUP}.')
        g = es.group('Destinations')
        g.quantity = es.n_dest

    @staticmethod
    def _check_actions_vs_groups(es: EnvSpec):
        need = {S.OP_DOORUSE: ('Doors', es.has_doors), S.OP_CLEAN: ('DirtPiles', es.has_dirt),
                S.OP_CHARGE: ('Batteries + ChargePods', es.has_batteries and es.n_pods > 0),
                S.OP_DEST: ('Destinations', es.n_dest > 0), S.OP_MACHINE: ('Machines', es.n_machines > 0),
                S.OP_ITEM: ('Items + Inventories + DropOffLocations',
                            es.n_items > 0 and es.has_inventories and es.n_dropoff > 0)}
        for a in es.agents:
            for act in a.actions:
                if act.opcode in need and not need[act.opcode][1]:
                    raise ValueError(f'{a.name}: action "{act.class_name}" needs Entities: {need[act.opcode][0]}.')

    # ------------------------------------------------------------------------------------ rules
    @staticmethod
    def _compile_rule(name: str, kw: dict, es: EnvSpec) -> RuleSpec:
        if name not in S.RULE_NAMES:
            raise NotImplementedError(f'Rule "{name}" is not supported by the B200 engine (built-in rules: '
                                      f'{sorted(S.RULE_NAMES)}).')
        op = S.RULE_NAMES[name]
        p = [0.0] * S.RULE_NPARAM

        def take(key, default):
            return kw.pop(key, default)

        if op == S.R_WATCH_COLLISIONS:          # environment/rules.py:258
            p[0] = float(take('reward', -0.5))
            p[1] = float(bool(take('done_at_collisions', False)))
            p[2] = float(take('reward_at_done', -1))
        elif op == S.R_RESPAWN_DIRT:            # clean_up/rules.py:30
            _require(es.has_dirt, name, 'DirtPiles')
            p[0] = float(int(take('respawn_freq', 15)))
            p[1] = float(int(take('respawn_n', 5)))
            p[2] = float(take('respawn_amount', 1.0))
            if p[2] == 0.0:
                raise NotImplementedError('RespawnDirt.respawn_amount == 0 is not supported.')
        elif op == S.R_SMEAR_DIRT:              # clean_up/rules.py:64 (never fires)
            _require(es.has_dirt, name, 'DirtPiles')
            if not float(take('smear_ratio', 0.2)) < 1:
                raise AssertionError("'Smear Amount' must be smaller than 1")
        elif op == S.R_DOOR_AUTO_CLOSE:         # doors/rules.py:10 (close_frequency unused)
            _require(es.has_doors, name, 'Doors')
            take('close_frequency', 10)
        elif op == S.R_DONE_ALL_DIRT:           # clean_up/rules.py:11
            _require(es.has_dirt, name, 'DirtPiles')
            p[0] = float(take('reward', 4.5))
        elif op in (S.R_BATTERY_DECHARGE, S.R_DONE_BATTERY):   # batteries/rules.py:11, 92
            _require(es.has_batteries, name, 'Batteries')
            take('initial_charge', 0.8)
            take('battery_charge_reward', 0.1)
            take('battery_failed_reward', -0.1)
            pac = take('per_action_costs', 0.02)
            if isinstance(pac, dict):
                # cost by the class name of the action the agent took this tick (batteries/rules.py:55-58); a paralysed agent's
                # default state is a valid 'Noop' (entity.py:19).  A missing name is a KeyError at the first tick there.
                es.act_costs = []
                for ag in es.agents:
                    names = [a.class_name for a in ag.actions] + ['Noop']
                    missing = [n for n in names if n not in pac]
                    if missing:
                        raise KeyError(f'{name}.per_action_costs has no entry for {missing[0]!r} ({ag.name}).')
                    es.act_costs.append([float(pac[n]) for n in names])
                p[5] = 1.0
            else:
                p[0] = float(pac)
            p[1] = float(take('battery_discharge_reward', -1.0))
            p[2] = float(bool(take('paralyze_agents_on_discharge', False)))
            if op == S.R_DONE_BATTERY:
                p[3] = float(take('reward_discharge_done', -1.0))
                # constants are cross-named in the reference: SINGLE = "grouped" is the default and the only
                # value for which the rule can fire (batteries/constants.py, rules.py:122-128)
                p[4] = float(take('mode', 'grouped') == 'grouped')
        elif op in (S.R_DEST_REACH_REWARD, S.R_DONE_DEST):     # destinations/rules.py:18, 58
            _require(es.n_dest > 0, name, 'Destinations')
            p[0] = float(take('dest_reach_reward', 1.0))
            if op == S.R_DONE_DEST:
                cond = take('condition', 'any')
                if cond not in ('any', 'all', 'simultaneous'):
                    raise AssertionError('condition must be one of any / all / simultaneous')
                p[1] = float(('any', 'all', 'simultaneous').index(cond))
                p[2] = float(take('reward_at_done', 5.0))
        elif op == S.R_RESPAWN_ITEMS:           # items/rules.py:11 (no state effect)
            _require(es.n_items > 0, name, 'Items')
            for k in ('n_items', 'respawn_freq', 'n_locations'):
                take(k, None)
        elif op == S.R_MOVE_MAINTAINERS:
            _require(es.n_maint > 0 and es.n_machines > 0 and es.has_doors, name, 'Maintainers + Machines + Doors')
        elif op == S.R_DONE_MAINT_COLLISION:
            _require(es.n_maint > 0, name, 'Maintainers')
        elif op == S.R_RANDOM_INITIAL_STEPS:    # environment/rules.py:328-355
            if 'random_steps' not in kw:
                raise TypeError("DoRandomInitialSteps.__init__() missing 1 required positional argument: 'random_steps'")
            es.random_initial_steps = int(take('random_steps', 0))
            p[0] = float(es.random_initial_steps)
        elif op == S.R_DONE_MAX_STEPS:          # environment/rules.py:204
            p[0] = float(int(take('max_steps', 500)))
            if not 0 <= p[0] <= 65535:
                raise ValueError('DoneAtMaxStepsReached.max_steps must be within 0..65535 (16-bit step counter).')
        if kw:
            raise TypeError(f'{name}.__init__() got unexpected keyword argument(s) {sorted(kw)}.')
        return RuleSpec(name, op, p)


def _pick(kw: dict, key: str, default: float) -> float:
    v = kw.pop(key, None)
    return float(default if v is None else v)


def _no_extra(kw: dict, allowed, action: str):
    extra = set(kw) - set(allowed)
    if extra:
        raise TypeError(f'{action}.__init__() got unexpected keyword argument(s) {sorted(extra)}.')


def _require(cond: bool, rule: str, group: str):
    if not cond:
        raise ValueError(f'Rule "{rule}" needs Entities: {group}.')


def named_action_space(es: EnvSpec) -> Dict[str, Dict[str, int]]:
    """agents.py:50-60: equal action names collapse in the dict, indices still count."""
    return {a.name: {act.name: i for i, act in enumerate(a.actions)} for a in es.agents}
