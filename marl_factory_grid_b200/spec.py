"""EnvSpec: the immutable, compiled form of (yaml config, level .txt).

This is the host-side product of the config compiler (config_parser.py) and the single input of
both the CUDA engine (packed into the C-ABI `MfgSpec`, include/mfg_b200.h) and the test oracle.
It replaces the reference's object graph set-up (`Factory.__init__`,
marl_factory_grid/environment/factory.py:81-129) by flat tables:

  * level:   wall bitmap, floor tile list, door tiles                     (utils/level_parser.py:26-102)
  * agents:  per-agent action table (opcode, direction, valid/fail reward) (utils/config_parser.py:128-199)
  * rules:   ordered rule program (yaml order, spawn rules appended)      (utils/config_parser.py:201-274)
  * obs:     per-agent observation channel program                        (utils/observation_builder.py:237-277)

All numeric ids below are mirrored verbatim in include/mfg_b200.h.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# ---------------------------------------------------------------------------------------------
# action opcodes (marl_factory_grid/environment/actions.py, modules/*/actions.py)
# ---------------------------------------------------------------------------------------------
OP_NOOP, OP_MOVE, OP_DOORUSE, OP_CLEAN, OP_ITEM, OP_CHARGE, OP_DEST, OP_MACHINE = range(8)

# direction ids in Move8 order (environment/actions.py:145-146) with MOVEMAP deltas (utils/helpers.py:36-43);
# x is the ROW, y the COLUMN.
DIR_NAMES = ['north', 'east', 'south', 'west', 'north_east', 'south_east', 'south_west', 'north_west']
DIR_CLASS = ['North', 'East', 'South', 'West', 'NorthEast', 'SouthEast', 'SouthWest', 'NorthWest']
DIR_DELTA = [(-1, 0), (0, 1), (1, 0), (0, -1), (-1, 1), (1, 1), (1, -1), (-1, -1)]

# maintainer tape codes (tests/golden/make_golden.py MAINT_CODE): 0..7 = move dir, then
MAINT_NOOP, MAINT_DOORUSE, MAINT_MACHINE = 8, 9, 10

# ---------------------------------------------------------------------------------------------
# entity groups with a position (ids index the per-agent channel-mask table of the obs program)
# ---------------------------------------------------------------------------------------------
(G_WALLS, G_DOORS, G_DIRT, G_ITEMS, G_DROPOFF, G_PODS, G_DEST, G_MACHINES, G_MAINT) = range(9)
N_GROUPS = 9
G_AGENT0 = N_GROUPS            # term id of agent j is G_AGENT0 + j
GROUP_NAMES = {
    'Walls': G_WALLS, 'Doors': G_DOORS, 'DirtPiles': G_DIRT, 'Items': G_ITEMS, 'DropOffLocations': G_DROPOFF,
    'ChargePods': G_PODS, 'Destinations': G_DEST, 'Machines': G_MACHINES, 'Maintainers': G_MAINT,
}
# observation encodings: doors/constants.py:10-11, machines/entitites.py:27, others 1
ENC_DOOR_OPEN, ENC_DOOR_CLOSED, ENC_MACHINE = 0.4444, 0.6666, 15.0

# channel kinds of the observation program
CH_TERMS, CH_ZERO, CH_BATTERY, CH_GLOBALPOS = range(4)

# ---------------------------------------------------------------------------------------------
# rule opcodes (environment/rules.py, modules/*/rules.py)
# ---------------------------------------------------------------------------------------------
(R_WATCH_COLLISIONS, R_RESPAWN_DIRT, R_SMEAR_DIRT, R_DOOR_AUTO_CLOSE, R_DONE_ALL_DIRT, R_BATTERY_DECHARGE,
 R_DONE_BATTERY, R_DEST_REACH_REWARD, R_DONE_DEST, R_RESPAWN_ITEMS, R_MOVE_MAINTAINERS, R_DONE_MAINT_COLLISION,
 R_DONE_MAX_STEPS, R_RANDOM_INITIAL_STEPS) = range(14)
RULE_NAMES = {
    'WatchCollisions': R_WATCH_COLLISIONS, 'RespawnDirt': R_RESPAWN_DIRT, 'EntitiesSmearDirtOnMove': R_SMEAR_DIRT,
    'DoorAutoClose': R_DOOR_AUTO_CLOSE, 'DoneOnAllDirtCleaned': R_DONE_ALL_DIRT,
    'BatteryDecharge': R_BATTERY_DECHARGE, 'DoneAtBatteryDischarge': R_DONE_BATTERY,
    'DestinationReachReward': R_DEST_REACH_REWARD, 'DoneAtDestinationReach': R_DONE_DEST,
    'RespawnItems': R_RESPAWN_ITEMS, 'MoveMaintainers': R_MOVE_MAINTAINERS,
    'DoneAtMaintainerCollision': R_DONE_MAINT_COLLISION, 'DoneAtMaxStepsReached': R_DONE_MAX_STEPS,
    'DoRandomInitialSteps': R_RANDOM_INITIAL_STEPS,
}
# how Destinations are spawned (modules/destinations/rules.py:95-162)
DEST_FREE, DEST_ON_AGENT, DEST_PER_AGENT = range(3)
RULE_NPARAM = 6                 # f64 parameters per rule entry (meaning depends on the opcode)

# capacities shared with the kernels (include/mfg_b200.h)
MAX_AGENTS = 16
MAX_ACTIONS = 32
MAX_DOORS = 64
MAX_DIRT = 64
MAX_RULES = 32
MAX_CHANNELS = 32               # per agent
MAX_SMALL_GROUP = 32            # items / pods / dests / drop-offs / machines / maintainers, each
NO_TILE = 0xFFFF

DOOR_AUTO_CLOSE_INTERVAL = 10   # modules/doors/entitites.py:69 (the rule's close_frequency is unused)
DIRT_PILE_MAX = 5.0             # modules/clean_up/entitites.py default max_local_amount of a DirtPile
CHARGE_RATE = 0.4               # modules/batteries/entitites.py:98


@dataclass
class ActionSpec:
    name: str                   # reference Action.name  (e.g. 'north', 'use_door')
    class_name: str             # reference class name    (info-dict key, e.g. 'North', 'DoorUse')
    opcode: int
    direction: int = 0
    valid_reward: float = 0.0
    fail_reward: float = 0.0
    aux_reward: float = 0.0     # ItemAction: failed drop-off reward


@dataclass
class ChannelSpec:
    name: str                   # layer name as in OBSBuilder.obs_layers
    kind: int
    terms: List[int] = field(default_factory=list)     # CH_TERMS: group ids / G_AGENT0 + j, in names order


@dataclass
class AgentSpec:
    name: str                   # 'Agent[Wolfgang]'
    actions: List[ActionSpec]
    channels: List[ChannelSpec]
    positions: List[Tuple[int, int]] = field(default_factory=list)
    is_blocking_pos: bool = False


@dataclass
class RuleSpec:
    name: str
    opcode: int
    params: List[float]


@dataclass
class GroupSpec:
    """One `Entities:` yaml entry, in yaml order (= spawn order, = uid listing order)."""
    name: str
    quantity: int = 0
    coords: Optional[List[Tuple[int, int]]] = None       # fixed coordinates instead of a quantity


@dataclass
class EnvSpec:
    # ---- level
    level_name: str
    H: int
    W: int
    walls: np.ndarray           # bool [H, W]
    floor: np.ndarray           # int32 [F, 2] row-major non-wall tiles (doors included)
    door_pos: np.ndarray        # int32 [ND, 2] row-major 'D' tiles ([] when Doors are not configured)
    # ---- general
    pomdp_r: int
    env_seed: int
    individual_rewards: bool
    # ---- agents / rules / entities
    agents: List[AgentSpec]
    rules: List[RuleSpec]
    groups: List[GroupSpec]
    # ---- module parameters
    has_doors: bool = False
    has_dirt: bool = False
    has_batteries: bool = False
    has_inventories: bool = False
    has_globalpos: bool = False
    dirt_quantity: int = 10
    dirt_initial_amount: float = 2.0
    dirt_clean_amount: float = 1.0
    dirt_max_global: float = 20.0
    dirt_n_var: float = 0.2
    dirt_amount_var: float = 0.2
    dirt_slots: int = 40
    battery_initial: float = 1.0
    n_items: int = 0
    n_dropoff: int = 0
    n_pods: int = 0
    n_dest: int = 0
    n_machines: int = 0
    n_maint: int = 0
    # bound destinations (SpawnDestinationOnAgent / SpawnDestinationsPerAgent) and DoRandomInitialSteps
    dest_mode: int = 0
    dest_bound: List[int] = field(default_factory=list)                     # agent index per destination (-1: unbound)
    dest_cands: List[List[Tuple[int, int]]] = field(default_factory=list)   # per-agent mode: candidate tiles ([] = any floor tile)
    random_initial_steps: int = 0
    act_costs: List[List[float]] = field(default_factory=list)              # BatteryDecharge.per_action_costs as a dict

    # ------------------------------------------------------------------ derived helpers
    @property
    def n_agents(self) -> int:
        return len(self.agents)

    @property
    def n_doors(self) -> int:
        return int(len(self.door_pos))

    @property
    def n_floor(self) -> int:
        return int(len(self.floor))

    @property
    def obs_d(self) -> int:
        """Window diameter of the egocentric POMDP observation (pomdp_r >= 1)."""
        return 2 * self.pomdp_r + 1

    @property
    def obs_shape(self) -> Tuple[int, int]:
        """observation_builder.py:51: (D, D) for POMDP, the whole level when pomdp_r == 0."""
        return (self.obs_d, self.obs_d) if self.pomdp_r else (self.H, self.W)

    @property
    def channels_per_agent(self) -> List[int]:
        return [len(a.channels) for a in self.agents]

    @property
    def channel_offsets(self) -> List[int]:
        off, out = 0, []
        for c in self.channels_per_agent:
            out.append(off)
            off += c
        return out

    @property
    def total_channels(self) -> int:
        return sum(self.channels_per_agent)

    @property
    def n_actions(self) -> List[int]:
        return [len(a.actions) for a in self.agents]

    def rule(self, opcode: int) -> Optional[RuleSpec]:
        return next((r for r in self.rules if r.opcode == opcode), None)

    def group(self, name: str) -> Optional[GroupSpec]:
        return next((g for g in self.groups if g.name == name), None)

    # ------------------------------------------------------------------ roofline bookkeeping
    def algorithmic_state_bytes(self) -> int:
        """Minimal SoA state bytes per env, SURVEY.md §8(d) formula (layout independent)."""
        A = self.n_agents
        s = 8 + 2 * A
        if self.has_batteries:
            s += 8 * A
        if self.has_doors:
            nd = self.n_doors
            s += nd + (nd + 7) // 8
        if self.has_dirt:
            s += 2 + 40 * 10
        if self.n_items:
            s += 2 * self.n_items + 2
        s += 2 * (self.n_dropoff + self.n_pods + self.n_dest + self.n_machines)
        if self.n_dest:
            s += 1
        if self.n_maint:
            s += 16
        return s

    def algorithmic_bytes_per_env_step(self) -> int:
        """B = obs f32 write + actions read + reward write + done + 2*S   (SURVEY.md §8(d))."""
        A = self.n_agents
        d2 = self.obs_d ** 2
        return 4 * d2 * self.total_channels + 4 * A + 4 * A + 1 + 2 * self.algorithmic_state_bytes()
