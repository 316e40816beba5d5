//! oxmpl-cuda — GPU backend behind oxmpl's existing traits.
//!
//! * [`GpuValidityChecker`] implements `oxmpl::base::validity::StateValidityChecker` for
//!   `RealVectorState` (declarative spheres / AABBs / capsules), so it drops into
//!   `Planner::setup(problem_def, Arc::new(checker))` of the stock planners unchanged.
//! * [`GpuRRT`], [`GpuRRTConnect`] and [`GpuRRTStar`] implement `oxmpl::base::planner::Planner`
//!   with the constructors and the `setup` / `solve(timeout)` contract of
//!   `oxmpl::geometric::{RRT, RRTConnect, RRTStar}` (rrt.rs:75, rrt_connect.rs:86, rrt_star.rs:65),
//!   but keep their trees in GPU-resident stores and run the loop as one persistent kernel per
//!   chunk of pre-drawn samples (`oxg_rrt_solve`, `oxg_rrt_connect_solve`, `oxg_rrt_star_solve`).
//! * [`GpuPRM`] is `oxmpl::geometric::PRM` (prm.rs:70-154, 227-307): `construct_roadmap` batches
//!   the incremental radius build through `oxg_prm_build_radius`, `solve` runs the start
//!   connection, the goal scan and the reference-order BFS on the device (`oxg_prm_solve`).
//! * [`GpuConeChecker`], [`GpuSo3RRT`] and [`GpuSo3RRTConnect`] cover `SO3StateSpace`
//!   (oxmpl/tests/{rrt,rrt_connect}_so3ss_tests.rs): cone validity on the device, and the two
//!   planner loops on `oxg_so3_rrt_solve` / `oxg_so3_rrt_connect_solve` (include/oxg_so3.h), which
//!   scan and motion-check a chunk of samples per launch and commit them in the reference's order.
//!
//! `Planner::setup` receives `Arc<dyn StateValidityChecker<S>>` like every reference planner.
//! The GPU planners need the checker's device handle, and a trait object cannot be downcast, so
//! [`GpuValidityChecker::into_arc`] records the address of the `Arc`'s payload in a process-wide
//! registry; `setup` looks the incoming trait object up by that address and panics — loudly, like
//! the reference's `assert_eq!`s — if it is some other checker (there is no CPU fallback).
//!
//! Sampling stays in Rust and uses exactly the reference's call sequence
//! (`rng.random_bool(goal_bias)` then `goal.sample_goal` or `space.sample_uniform`,
//! rrt.rs:177-184), which does not depend on the tree — so drawing a chunk up front changes
//! nothing the planner can observe.
pub mod ffi;

use std::{
    collections::HashMap,
    ffi::CStr,
    sync::{Arc, Mutex, OnceLock},
    time::{Duration, Instant},
};

use oxmpl::base::{
    error::PlanningError,
    goal::{Goal, GoalSampleableRegion},
    planner::{Path, Planner},
    problem_definition::ProblemDefinition,
    space::{RealVectorStateSpace, StateSpace},
    state::RealVectorState,
    validity::StateValidityChecker,
};
use rand::Rng;

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::oxg_last_error()).to_string_lossy().into_owned() }
}
fn check(code: i32) {
    // the reference's seams have no error channel (is_valid returns bool); a backend failure is
    // a programming / environment error and panics like the reference's assert_eq! do
    assert!(code == ffi::OXG_OK, "liboxg error {code}: {}", last_error());
}

/// Declarative obstacle set: `impl StateValidityChecker<RealVectorState>`.
pub struct GpuValidityChecker {
    world: ffi::oxg_world,
    dim: u32,
}
// The reference's planners hold `Arc<dyn StateValidityChecker<S>>` without Send/Sync bounds and
// are single-threaded; a handle must be used from one thread at a time (include/oxg.h).
impl GpuValidityChecker {
    /// spheres: rows `[c.., r]`; aabbs: rows `[lo.., hi..]`; capsules: rows `[a.., b.., r]`.
    pub fn new(dim: usize, spheres: &[f64], aabbs: &[f64], capsules: &[f64]) -> Self {
        let d = dim as u64;
        let mut world = std::ptr::null_mut();
        check(unsafe {
            ffi::oxg_world_create(dim as u32, spheres.as_ptr(), spheres.len() as u64 / (d + 1),
                                  aabbs.as_ptr(), aabbs.len() as u64 / (2 * d), capsules.as_ptr(),
                                  capsules.len() as u64 / (2 * d + 1), &mut world)
        });
        Self { world, dim: dim as u32 }
    }
    fn space(&self) -> ffi::oxg_space_desc {
        ffi::oxg_space_desc { n_real: self.dim, n_so2: 0, w_real: 1.0, w_so2: 1.0 }
    }
    pub(crate) fn handle(&self) -> ffi::oxg_world { self.world }

    /// Wraps the checker for `Planner::setup` and registers it so the GPU planners can recover the
    /// device handle from the `Arc<dyn StateValidityChecker<_>>` the reference's signature passes.
    pub fn into_arc(self) -> Arc<Self> {
        let arc = Arc::new(self);
        registry().lock().unwrap().insert(Arc::as_ptr(&arc) as *const u8 as usize, arc.world as usize);
        arc
    }
}
fn registry() -> &'static Mutex<HashMap<usize, usize>> {
    static R: OnceLock<Mutex<HashMap<usize, usize>>> = OnceLock::new();
    R.get_or_init(|| Mutex::new(HashMap::new()))
}
/// The device handle behind a trait object handed to `setup`, or a panic naming the problem.
fn gpu_world_of(vc: &Arc<dyn StateValidityChecker<RealVectorState>>) -> ffi::oxg_world {
    let key = Arc::as_ptr(vc) as *const u8 as usize;
    match registry().lock().unwrap().get(&key) {
        Some(w) => *w as ffi::oxg_world,
        None => panic!(
            "oxmpl-cuda planners need a GpuValidityChecker created with into_arc(): the checker \
             passed to setup() is not a registered declarative GPU checker (no CPU fallback)"
        ),
    }
}
impl Drop for GpuValidityChecker {
    fn drop(&mut self) {
        registry().lock().unwrap().retain(|_, w| *w != self.world as usize);
        unsafe { ffi::oxg_world_destroy(self.world); }
    }
}
impl StateValidityChecker<RealVectorState> for GpuValidityChecker {
    /// validity.rs:47 — one state per call (correct but slow: the GPU planners below use the
    /// batched / fused paths instead of calling this per interpolated state).
    fn is_valid(&self, state: &RealVectorState) -> bool {
        let mut out = 0u8;
        let sp = self.space();
        check(unsafe { ffi::oxg_is_valid(self.world, &sp, state.values.as_ptr(), 1, &mut out) });
        out != 0
    }
}

/// A goal the fused kernels can test on the device: `distance(state, target) <= radius`
/// (the CircularGoalRegion of the reference's tests).  Any other `Goal` works with the unfused
/// per-iteration calls (`oxg_nn1` + `oxg_check_motion`).
pub trait BallGoal: Goal<RealVectorState> + GoalSampleableRegion<RealVectorState> {
    fn target(&self) -> &RealVectorState;
    fn radius(&self) -> f64;
}

struct Store(ffi::oxg_store, usize);
impl Store {
    fn new(dim: usize) -> Self {
        let sp = ffi::oxg_space_desc { n_real: dim as u32, n_so2: 0, w_real: 1.0, w_so2: 1.0 };
        let mut s = std::ptr::null_mut();
        check(unsafe { ffi::oxg_store_create(&sp, 1024, &mut s) });
        Store(s, dim)
    }
    fn push(&self, st: &RealVectorState) {
        let mut first = 0u64;
        check(unsafe { ffi::oxg_store_append(self.0, st.values.as_ptr(), 1, &mut first) });
    }
    fn get(&self, i: usize) -> RealVectorState {
        let mut v = vec![0.0; self.1];
        check(unsafe { ffi::oxg_store_read(self.0, i as u64, 1, v.as_mut_ptr()) });
        RealVectorState { values: v }
    }
}
impl Drop for Store {
    fn drop(&mut self) { unsafe { ffi::oxg_store_destroy(self.0); } }
}

/// `oxmpl::geometric::RRT` with the tree on the GPU.  Same fields, constructor and trait.
pub struct GpuRRT<G: BallGoal> {
    pub max_distance: f64,
    pub goal_bias: f64,
    /// iterations per kernel launch
    pub chunk: usize,
    /// tree capacity (the reference is bounded by the timeout only)
    pub max_nodes: usize,
    problem_def: Option<Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>>,
    checker: Option<(Arc<dyn StateValidityChecker<RealVectorState>>, ffi::oxg_world)>,
    tree: Option<Store>,
}

impl<G: BallGoal> GpuRRT<G> {
    pub fn new(max_distance: f64, goal_bias: f64) -> Self {
        Self { max_distance, goal_bias, chunk: 1 << 14, max_nodes: 1 << 20, problem_def: None,
               checker: None, tree: None }
    }
}

/// `chunk` iterations of the reference's per-iteration draw (rrt.rs:177-184), row-major.
fn draw_stream<G: BallGoal>(
    pd: &ProblemDefinition<RealVectorState, RealVectorStateSpace, G>, goal_bias: f64, chunk: usize,
    rng: &mut impl Rng,
) -> Vec<f64> {
    let mut stream = Vec::with_capacity(chunk * pd.space.dimension);
    for _ in 0..chunk {
        let q = if rng.random_bool(goal_bias) {
            pd.goal.sample_goal(rng).unwrap()
        } else {
            pd.space.sample_uniform(rng).unwrap()
        };
        stream.extend_from_slice(&q.values);
    }
    stream
}

/// reconstruct_path (rrt.rs:118-128): walk the parents from `last` to the root, reversed.
fn chain(tree: &Store, parents: &[i64], last: i64) -> Vec<RealVectorState> {
    let mut states = Vec::new();
    let mut i = last;
    while i >= 0 {
        states.push(tree.get(i as usize));
        i = parents[i as usize];
    }
    states.reverse();
    states
}

impl<G: BallGoal> Planner<RealVectorState, RealVectorStateSpace, G> for GpuRRT<G> {
    fn setup(
        &mut self,
        problem_def: Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>,
        validity_checker: Arc<dyn StateValidityChecker<RealVectorState>>,
    ) {
        // rrt.rs:139-156: remember the problem and the checker, clear the tree, push the start state
        let world = gpu_world_of(&validity_checker);
        self.checker = Some((validity_checker, world));
        let tree = Store::new(problem_def.space.dimension);
        tree.push(&problem_def.start_states[0]);
        self.tree = Some(tree);
        self.problem_def = Some(problem_def);
    }

    fn solve(&mut self, timeout: Duration) -> Result<Path<RealVectorState>, PlanningError> {
        let pd = self.problem_def.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let world = self.checker.as_ref().ok_or(PlanningError::PlannerUninitialised)?.1;
        let tree = self.tree.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let start_time = Instant::now();
        let mut rng = rand::rng();
        let mut parents = vec![-1i64; self.max_nodes];
        let mut prm = ffi::oxg_plan_params {
            max_distance: self.max_distance,
            resolution: pd.space.get_longest_valid_segment_length() * 0.1, // rrt.rs:97
            goal_radius: pd.goal.radius(),
            max_nodes: self.max_nodes as u64,
            stop_at_goal: 1,
            ..Default::default()
        };
        loop {
            let elapsed = start_time.elapsed();
            if elapsed > timeout {
                return Err(PlanningError::Timeout);
            }
            prm.timeout_s = (timeout - elapsed).as_secs_f64();
            // the reference's per-iteration draw (rrt.rs:177-184), `chunk` iterations ahead
            let stream = draw_stream(pd, self.goal_bias, self.chunk, &mut rng);
            let mut r = ffi::oxg_plan_result::default();
            check(unsafe {
                ffi::oxg_rrt_solve(world, tree.0, &prm, pd.goal.target().values.as_ptr(),
                                   stream.as_ptr(), self.chunk as u64, parents.as_mut_ptr(), &mut r)
            });
            if r.solved != 0 {
                return Ok(Path(chain(tree, &parents, r.goal_node[0])));
            }
            if r.status == 3 || r.status == 2 {
                return Err(PlanningError::Timeout);
            }
        }
    }
}

/// `oxmpl::geometric::RRTConnect` (rrt_connect.rs:63-310) with both trees on the GPU.
pub struct GpuRRTConnect<G: BallGoal> {
    pub max_distance: f64,
    pub goal_bias: f64,
    pub chunk: usize,
    pub max_nodes: usize,
    problem_def: Option<Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>>,
    checker: Option<(Arc<dyn StateValidityChecker<RealVectorState>>, ffi::oxg_world)>,
    start_tree: Option<Store>,
    goal_tree: Option<Store>,
}

impl<G: BallGoal> GpuRRTConnect<G> {
    /// rrt_connect.rs:86
    pub fn new(max_distance: f64, goal_bias: f64) -> Self {
        Self { max_distance, goal_bias, chunk: 1 << 14, max_nodes: 1 << 20, problem_def: None,
               checker: None, start_tree: None, goal_tree: None }
    }
}

impl<G: BallGoal> Planner<RealVectorState, RealVectorStateSpace, G> for GpuRRTConnect<G> {
    fn setup(
        &mut self,
        problem_def: Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>,
        validity_checker: Arc<dyn StateValidityChecker<RealVectorState>>,
    ) {
        // rrt_connect.rs:199-225: the start tree holds the start state, the goal tree one sampled goal
        let world = gpu_world_of(&validity_checker);
        self.checker = Some((validity_checker, world));
        let dim = problem_def.space.dimension;
        let (a, b) = (Store::new(dim), Store::new(dim));
        a.push(&problem_def.start_states[0]);
        let mut rng = rand::rng();
        b.push(&problem_def.goal.sample_goal(&mut rng).unwrap());
        self.start_tree = Some(a);
        self.goal_tree = Some(b);
        self.problem_def = Some(problem_def);
    }

    fn solve(&mut self, timeout: Duration) -> Result<Path<RealVectorState>, PlanningError> {
        let pd = self.problem_def.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let world = self.checker.as_ref().ok_or(PlanningError::PlannerUninitialised)?.1;
        let (ta, tb) = (self.start_tree.as_ref().unwrap(), self.goal_tree.as_ref().unwrap());
        let start_time = Instant::now();
        let mut rng = rand::rng();
        let (mut pa, mut pb) = (vec![-1i64; self.max_nodes], vec![-1i64; self.max_nodes]);
        let mut prm = ffi::oxg_plan_params {
            max_distance: self.max_distance,
            resolution: pd.space.get_longest_valid_segment_length() * 0.1, // rrt_connect.rs:173
            goal_radius: pd.goal.radius(),
            max_nodes: self.max_nodes as u64,
            stop_at_goal: 1,
            ..Default::default()
        };
        loop {
            let elapsed = start_time.elapsed();
            if elapsed > timeout {
                return Err(PlanningError::Timeout); // rrt_connect.rs:243-245
            }
            prm.timeout_s = (timeout - elapsed).as_secs_f64();
            let stream = draw_stream(pd, self.goal_bias, self.chunk, &mut rng); // rrt_connect.rs:258-262
            let mut r = ffi::oxg_plan_result::default();
            check(unsafe {
                ffi::oxg_rrt_connect_solve(world, ta.0, tb.0, &prm, pd.goal.target().values.as_ptr(),
                                           stream.as_ptr(), self.chunk as u64, pa.as_mut_ptr(),
                                           pb.as_mut_ptr(), &mut r)
            });
            if r.solved != 0 {
                // rrt_connect.rs:271-304: the start tree reached the goal itself, or the trees met
                let mut path = chain(ta, &pa, r.goal_node[0]);
                if r.goal_node[1] >= 0 {
                    let mut goal_side = chain(tb, &pb, r.goal_node[1]);
                    goal_side.reverse();
                    path.extend(goal_side.into_iter().skip(1)); // skip the duplicate connection point
                }
                return Ok(Path(path));
            }
            if r.status == 3 || r.status == 2 {
                return Err(PlanningError::Timeout);
            }
        }
    }
}

/// `oxmpl::geometric::RRTStar` (rrt_star.rs:40-291): radius neighbours, choose-parent, rewire.
pub struct GpuRRTStar<G: BallGoal> {
    pub max_distance: f64,
    pub goal_bias: f64,
    pub search_radius: f64,
    pub chunk: usize,
    pub max_nodes: usize,
    problem_def: Option<Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>>,
    checker: Option<(Arc<dyn StateValidityChecker<RealVectorState>>, ffi::oxg_world)>,
    tree: Option<Store>,
}

impl<G: BallGoal> GpuRRTStar<G> {
    /// rrt_star.rs:65
    pub fn new(max_distance: f64, goal_bias: f64, search_radius: f64) -> Self {
        Self { max_distance, goal_bias, search_radius, chunk: 1 << 14, max_nodes: 1 << 20,
               problem_def: None, checker: None, tree: None }
    }
}

impl<G: BallGoal> Planner<RealVectorState, RealVectorStateSpace, G> for GpuRRTStar<G> {
    fn setup(
        &mut self,
        problem_def: Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>,
        validity_checker: Arc<dyn StateValidityChecker<RealVectorState>>,
    ) {
        // rrt_star.rs:152-170: the tree holds the start state at cost 0
        let world = gpu_world_of(&validity_checker);
        self.checker = Some((validity_checker, world));
        let tree = Store::new(problem_def.space.dimension);
        tree.push(&problem_def.start_states[0]);
        self.tree = Some(tree);
        self.problem_def = Some(problem_def);
    }

    fn solve(&mut self, timeout: Duration) -> Result<Path<RealVectorState>, PlanningError> {
        let pd = self.problem_def.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let world = self.checker.as_ref().ok_or(PlanningError::PlannerUninitialised)?.1;
        let tree = self.tree.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let start_time = Instant::now();
        let mut rng = rand::rng();
        let mut parents = vec![-1i64; self.max_nodes];
        let mut costs = vec![0.0f64; self.max_nodes];
        let mut prm = ffi::oxg_plan_params {
            max_distance: self.max_distance,
            search_radius: self.search_radius,
            resolution: pd.space.get_longest_valid_segment_length() * 0.1, // rrt_star.rs:83
            goal_radius: pd.goal.radius(),
            max_nodes: self.max_nodes as u64,
            stop_at_goal: 1, // rrt_star.rs:285-288 returns at the first node that satisfies the goal
            ..Default::default()
        };
        loop {
            let elapsed = start_time.elapsed();
            if elapsed > timeout {
                return Err(PlanningError::Timeout); // rrt_star.rs:184-186
            }
            prm.timeout_s = (timeout - elapsed).as_secs_f64();
            let stream = draw_stream(pd, self.goal_bias, self.chunk, &mut rng);
            let mut r = ffi::oxg_plan_result::default();
            let rc = unsafe {
                ffi::oxg_rrt_star_solve(world, tree.0, &prm, pd.goal.target().values.as_ptr(),
                                        stream.as_ptr(), self.chunk as u64, parents.as_mut_ptr(),
                                        costs.as_mut_ptr(), &mut r)
            };
            if rc == ffi::OXG_E_CAPACITY {
                // more than OXG_RRT_STAR_MAX_NEIGHBOURS nodes inside search_radius (include/oxg.h)
                return Err(PlanningError::NoSolutionFound);
            }
            check(rc);
            if r.solved != 0 {
                return Ok(Path(chain(tree, &parents, r.goal_node[0])));
            }
            if r.status == 3 || r.status == 2 {
                return Err(PlanningError::Timeout);
            }
        }
    }
}

/// A roadmap node as `PRM::get_roadmap` returns it (prm.rs:24-29, 82-84).
#[derive(Clone)]
pub struct GpuPrmNode {
    pub state: RealVectorState,
    pub edges: Vec<usize>,
}

/// `oxmpl::geometric::PRM` (prm.rs:48-308): the roadmap's states live in a GPU store, its edges in
/// the lower-index CSR the library emits (`new_node.edges` at push time, prm.rs:135).
pub struct GpuPRM<G: BallGoal> {
    pub timeout: f64,
    pub connection_radius: f64,
    /// samples drawn per batch of the incremental build
    pub batch: usize,
    pub max_nodes: usize,
    problem_def: Option<Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>>,
    checker: Option<(Arc<dyn StateValidityChecker<RealVectorState>>, ffi::oxg_world)>,
    store: Option<Store>,
    lower_off: Vec<u64>,
    lower_idx: Vec<u32>,
}

impl<G: BallGoal> GpuPRM<G> {
    /// prm.rs:70
    pub fn new(timeout: f64, connection_radius: f64) -> Self {
        Self { timeout, connection_radius, batch: 1 << 13, max_nodes: 1 << 22, problem_def: None,
               checker: None, store: None, lower_off: vec![0], lower_idx: Vec::new() }
    }
    fn n_nodes(&self) -> usize { self.lower_off.len() - 1 }

    /// prm.rs:88-90
    pub fn set_problem_definition(
        &mut self, pd: Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>,
    ) {
        self.problem_def = Some(pd);
    }

    /// prm.rs:82-84: states plus symmetric edge lists (lower neighbours first, then the later nodes
    /// that connected to this one, both ascending — the order the reference's lists grow in).
    pub fn get_roadmap(&self) -> Vec<GpuPrmNode> {
        let n = self.n_nodes();
        let store = match self.store.as_ref() { Some(s) => s, None => return Vec::new() };
        let mut nodes: Vec<GpuPrmNode> =
            (0..n).map(|i| GpuPrmNode { state: store.get(i), edges: Vec::new() }).collect();
        for i in 0..n {
            for t in self.lower_off[i] as usize..self.lower_off[i + 1] as usize {
                let j = self.lower_idx[t] as usize;
                nodes[i].edges.push(j);
            }
        }
        for i in 0..n {
            for t in self.lower_off[i] as usize..self.lower_off[i + 1] as usize {
                nodes[self.lower_idx[t] as usize].edges.push(i); // prm.rs:143-145
            }
        }
        nodes
    }

    /// prm.rs:96-154, in batches: draw `batch` samples, keep the valid ones (prm.rs:123), append
    /// them and let the library connect each to the earlier nodes within `connection_radius` whose
    /// motion is valid (prm.rs:131-138) — identical to the one-at-a-time loop, because node i only
    /// ever looks at nodes j < i.
    pub fn construct_roadmap(&mut self) -> Result<(), PlanningError> {
        if self.n_nodes() > 0 {
            return Ok(()); // prm.rs:106-113: an existing roadmap is reused
        }
        let pd = self.problem_def.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let world = self.checker.as_ref().ok_or(PlanningError::PlannerUninitialised)?.1;
        let store = self.store.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let dim = pd.space.dimension;
        let sp = ffi::oxg_space_desc { n_real: dim as u32, n_so2: 0, w_real: 1.0, w_so2: 1.0 };
        let resolution = pd.space.get_longest_valid_segment_length() * 0.1; // prm.rs:168
        let start_time = Instant::now();
        let mut rng = rand::rng(); // prm.rs:115
        while start_time.elapsed().as_secs_f64() <= self.timeout && self.n_nodes() < self.max_nodes {
            let mut rows = Vec::with_capacity(self.batch * dim);
            for _ in 0..self.batch {
                rows.extend_from_slice(&pd.space.sample_uniform(&mut rng).unwrap().values); // prm.rs:122
            }
            let mut ok = vec![0u8; self.batch];
            check(unsafe { ffi::oxg_is_valid(world, &sp, rows.as_ptr(), self.batch as u64, ok.as_mut_ptr()) });
            let keep: Vec<f64> = rows.chunks(dim).zip(&ok).filter(|(_, v)| **v != 0)
                .flat_map(|(r, _)| r.iter().copied()).collect();
            let nk = keep.len() / dim;
            if nk == 0 {
                continue;
            }
            let mut first = 0u64;
            check(unsafe { ffi::oxg_store_append(store.0, keep.as_ptr(), nk as u64, &mut first) });
            let mut off = vec![0u64; nk + 1];
            let mut idx = vec![0u32; 64 * nk];
            let mut st = ffi::oxg_prm_stats::default();
            loop {
                let rc = unsafe {
                    ffi::oxg_prm_build_radius(world, store.0, self.connection_radius, resolution, first,
                                              first + nk as u64, off.as_mut_ptr(), idx.as_mut_ptr(),
                                              idx.len() as u64, &mut st)
                };
                if rc == ffi::OXG_E_CAPACITY {
                    idx.resize(off[nk] as usize, 0);
                    continue;
                }
                check(rc);
                break;
            }
            let base = *self.lower_off.last().unwrap();
            self.lower_off.extend(off[1..].iter().map(|o| base + o));
            self.lower_idx.extend_from_slice(&idx[..off[nk] as usize]);
        }
        Ok(())
    }
}

impl<G: BallGoal> Planner<RealVectorState, RealVectorStateSpace, G> for GpuPRM<G> {
    fn setup(
        &mut self,
        problem_def: Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>,
        validity_checker: Arc<dyn StateValidityChecker<RealVectorState>>,
    ) {
        // prm.rs:217-225: remember the problem and the checker, clear the roadmap
        let world = gpu_world_of(&validity_checker);
        self.checker = Some((validity_checker, world));
        self.store = Some(Store::new(problem_def.space.dimension));
        self.lower_off = vec![0];
        self.lower_idx.clear();
        self.problem_def = Some(problem_def);
    }

    fn solve(&mut self, _timeout: Duration) -> Result<Path<RealVectorState>, PlanningError> {
        let pd = self.problem_def.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let world = self.checker.as_ref().ok_or(PlanningError::PlannerUninitialised)?.1;
        let store = self.store.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        if self.n_nodes() == 0 {
            return Err(PlanningError::UnsampledStateSpace); // prm.rs:239-241
        }
        let start = &pd.start_states[0];
        let mut path = vec![0u32; self.n_nodes()];
        let mut r = ffi::oxg_prm_query_result::default();
        check(unsafe {
            ffi::oxg_prm_solve(world, store.0, self.lower_off.as_ptr(), self.lower_idx.as_ptr(),
                               self.connection_radius, pd.space.get_longest_valid_segment_length() * 0.1,
                               start.values.as_ptr(), pd.goal.target().values.as_ptr(), pd.goal.radius(),
                               path.as_mut_ptr(), path.len() as u32, &mut r)
        });
        match r.status {
            0 => {
                // reconstruct_path (prm.rs:189-208): the start state, then the roadmap nodes
                let mut states = vec![start.clone()];
                states.extend(path[..r.path_len as usize].iter().map(|i| store.get(*i as usize)));
                Ok(Path(states))
            }
            2 => Err(PlanningError::InvalidStartState), // prm.rs:243-246
            3 => Err(PlanningError::UnsampledStateSpace),
            _ => Err(PlanningError::NoSolutionFound), // prm.rs:266-268, 304
        }
    }
}

// ------------------------------------------------------------------------------------------
// SO(3): the forbidden-cone checker of oxmpl/tests/rrt_so3ss_tests.rs:46-56 on the GPU.
// `SO3StateSpace::distance` needs libm's acos, which has no bit-identical device routine; the
// library compares a = |q1 . q2| against thresholds found with this process's own libm, so the
// booleans are the reference's (include/oxg_so3.h).
// ------------------------------------------------------------------------------------------
pub struct GpuConeChecker {
    cones: ffi::oxg_so3_cones,
}
impl GpuConeChecker {
    /// `cones`: (centre, angle) pairs; a state is valid iff `distance(centre, state) > angle` for all.
    pub fn new(cones: &[(oxmpl::base::state::SO3State, f64)]) -> Self {
        let rows: Vec<f64> = cones.iter().flat_map(|(c, a)| [c.x, c.y, c.z, c.w, *a]).collect();
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::oxg_so3_cones_create(rows.as_ptr(), cones.len() as u64, &mut h) });
        Self { cones: h }
    }
    /// check_motion (rrt.rs:90-116) for one edge; `resolution` = lvsl * 0.1 computed by the caller.
    pub fn check_motion(&self, from: &oxmpl::base::state::SO3State, to: &oxmpl::base::state::SO3State,
                        resolution: f64) -> bool {
        let (a, b) = ([from.x, from.y, from.z, from.w], [to.x, to.y, to.z, to.w]);
        let mut ok = 0u8;
        check(unsafe {
            ffi::oxg_so3_check_motion(self.cones, a.as_ptr(), b.as_ptr(), 1, resolution, &mut ok,
                                      std::ptr::null_mut())
        });
        ok != 0
    }
}
impl Drop for GpuConeChecker {
    fn drop(&mut self) {
        registry().lock().unwrap().retain(|_, c| *c != self.cones as usize);
        unsafe { ffi::oxg_so3_cones_destroy(self.cones); }
    }
}
impl StateValidityChecker<oxmpl::base::state::SO3State> for GpuConeChecker {
    fn is_valid(&self, state: &oxmpl::base::state::SO3State) -> bool {
        let v = [state.x, state.y, state.z, state.w];
        let mut out = 0u8;
        check(unsafe { ffi::oxg_so3_is_valid(self.cones, v.as_ptr(), 1, &mut out) });
        out != 0
    }
}

impl GpuConeChecker {
    /// Wraps the checker for `Planner::setup` and registers its device handle (see
    /// `GpuValidityChecker::into_arc`).
    pub fn into_arc(self) -> Arc<Self> {
        let arc = Arc::new(self);
        registry().lock().unwrap().insert(Arc::as_ptr(&arc) as *const u8 as usize, arc.cones as usize);
        arc
    }
}
fn gpu_cones_of(vc: &Arc<dyn StateValidityChecker<oxmpl::base::state::SO3State>>) -> ffi::oxg_so3_cones {
    let key = Arc::as_ptr(vc) as *const u8 as usize;
    match registry().lock().unwrap().get(&key) {
        Some(c) => *c as ffi::oxg_so3_cones,
        None => panic!(
            "oxmpl-cuda SO3 planners need a GpuConeChecker created with into_arc() (no CPU fallback)"
        ),
    }
}

/// The ball goal of oxmpl/tests/rrt_so3ss_tests.rs:58-68 (`distance(state, target) <= radius`).
pub trait So3BallGoal: Goal<oxmpl::base::state::SO3State>
    + GoalSampleableRegion<oxmpl::base::state::SO3State>
{
    fn target(&self) -> &oxmpl::base::state::SO3State;
    fn radius(&self) -> f64;
}

type So3 = oxmpl::base::state::SO3State;
type So3Space = oxmpl::base::space::SO3StateSpace;

fn so3_row(s: &So3) -> [f64; 4] { [s.x, s.y, s.z, s.w] }

struct So3Store(ffi::oxg_store);
impl So3Store {
    fn new(root: &So3) -> Self {
        let sp = ffi::oxg_space_desc { n_real: 4, n_so2: 0, w_real: 1.0, w_so2: 1.0 };
        let mut s = std::ptr::null_mut();
        check(unsafe { ffi::oxg_store_create(&sp, 1024, &mut s) });
        let (row, mut first) = (so3_row(root), 0u64);
        check(unsafe { ffi::oxg_store_append(s, row.as_ptr(), 1, &mut first) });
        So3Store(s)
    }
    fn get(&self, i: usize) -> So3 {
        let mut v = [0.0f64; 4];
        check(unsafe { ffi::oxg_store_read(self.0, i as u64, 1, v.as_mut_ptr()) });
        So3 { x: v[0], y: v[1], z: v[2], w: v[3] }
    }
    fn chain(&self, parents: &[i64], last: i64) -> Vec<So3> {
        let (mut states, mut i) = (Vec::new(), last);
        while i >= 0 {
            states.push(self.get(i as usize));
            i = parents[i as usize];
        }
        states.reverse();
        states
    }
}
impl Drop for So3Store {
    fn drop(&mut self) { unsafe { ffi::oxg_store_destroy(self.0); } }
}

fn so3_stream<G: So3BallGoal>(
    pd: &ProblemDefinition<So3, So3Space, G>, goal_bias: f64, chunk: usize, rng: &mut impl Rng,
) -> Vec<f64> {
    let mut stream = Vec::with_capacity(chunk * 4);
    for _ in 0..chunk {
        let q = if rng.random_bool(goal_bias) {
            pd.goal.sample_goal(rng).unwrap()
        } else {
            pd.space.sample_uniform(rng).unwrap()
        };
        stream.extend_from_slice(&so3_row(&q));
    }
    stream
}

/// `oxmpl::geometric::RRT` over `SO3StateSpace` (rrt.rs:63-227; oxmpl/tests/rrt_so3ss_tests.rs):
/// `oxg_so3_rrt_solve` consumes a pre-drawn stream in speculative chunks and returns the tree the
/// one-sample-at-a-time loop builds.
pub struct GpuSo3RRT<G: So3BallGoal> {
    pub max_distance: f64,
    pub goal_bias: f64,
    pub chunk: usize,
    pub max_nodes: usize,
    problem_def: Option<Arc<ProblemDefinition<So3, So3Space, G>>>,
    cones: Option<(Arc<dyn StateValidityChecker<So3>>, ffi::oxg_so3_cones)>,
    tree: Option<So3Store>,
}
impl<G: So3BallGoal> GpuSo3RRT<G> {
    pub fn new(max_distance: f64, goal_bias: f64) -> Self {
        Self { max_distance, goal_bias, chunk: 1 << 12, max_nodes: 1 << 20, problem_def: None,
               cones: None, tree: None }
    }
}
impl<G: So3BallGoal> Planner<So3, So3Space, G> for GpuSo3RRT<G> {
    fn setup(&mut self, problem_def: Arc<ProblemDefinition<So3, So3Space, G>>,
             validity_checker: Arc<dyn StateValidityChecker<So3>>) {
        let cones = gpu_cones_of(&validity_checker);
        self.cones = Some((validity_checker, cones));
        self.tree = Some(So3Store::new(&problem_def.start_states[0])); // rrt.rs:147-155
        self.problem_def = Some(problem_def);
    }

    fn solve(&mut self, timeout: Duration) -> Result<Path<So3>, PlanningError> {
        let pd = self.problem_def.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let cones = self.cones.as_ref().ok_or(PlanningError::PlannerUninitialised)?.1;
        let tree = self.tree.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let (start_time, mut rng) = (Instant::now(), rand::rng());
        let mut parents = vec![-1i64; self.max_nodes];
        let mut prm = ffi::oxg_plan_params {
            max_distance: self.max_distance,
            resolution: pd.space.get_longest_valid_segment_length() * 0.1, // rrt.rs:97
            goal_radius: pd.goal.radius(),
            max_nodes: self.max_nodes as u64,
            stop_at_goal: 1,
            ..Default::default()
        };
        let target = so3_row(pd.goal.target());
        loop {
            let elapsed = start_time.elapsed();
            if elapsed > timeout {
                return Err(PlanningError::Timeout);
            }
            prm.timeout_s = (timeout - elapsed).as_secs_f64();
            let stream = so3_stream(pd, self.goal_bias, self.chunk, &mut rng);
            let mut r = ffi::oxg_plan_result::default();
            check(unsafe {
                ffi::oxg_so3_rrt_solve(tree.0, cones, &prm, target.as_ptr(), stream.as_ptr(),
                                       self.chunk as u64, parents.as_mut_ptr(), &mut r)
            });
            if r.solved != 0 {
                return Ok(Path(tree.chain(&parents, r.goal_node[0])));
            }
            if r.status == 3 || r.status == 2 {
                return Err(PlanningError::Timeout);
            }
        }
    }
}

/// `oxmpl::geometric::RRTConnect` over `SO3StateSpace` (rrt_connect.rs:63-310;
/// oxmpl/tests/rrt_connect_so3ss_tests.rs) on `oxg_so3_rrt_connect_solve`.
pub struct GpuSo3RRTConnect<G: So3BallGoal> {
    pub max_distance: f64,
    pub goal_bias: f64,
    pub chunk: usize,
    pub max_nodes: usize,
    problem_def: Option<Arc<ProblemDefinition<So3, So3Space, G>>>,
    cones: Option<(Arc<dyn StateValidityChecker<So3>>, ffi::oxg_so3_cones)>,
    start_tree: Option<So3Store>,
    goal_tree: Option<So3Store>,
}
impl<G: So3BallGoal> GpuSo3RRTConnect<G> {
    pub fn new(max_distance: f64, goal_bias: f64) -> Self {
        Self { max_distance, goal_bias, chunk: 1 << 12, max_nodes: 1 << 20, problem_def: None,
               cones: None, start_tree: None, goal_tree: None }
    }
}
impl<G: So3BallGoal> Planner<So3, So3Space, G> for GpuSo3RRTConnect<G> {
    fn setup(&mut self, problem_def: Arc<ProblemDefinition<So3, So3Space, G>>,
             validity_checker: Arc<dyn StateValidityChecker<So3>>) {
        let cones = gpu_cones_of(&validity_checker);
        self.cones = Some((validity_checker, cones));
        self.start_tree = Some(So3Store::new(&problem_def.start_states[0]));
        // rrt_connect.rs:201-216: the goal tree is rooted at one sampled goal state
        let goal_root = problem_def.goal.sample_goal(&mut rand::rng()).unwrap();
        self.goal_tree = Some(So3Store::new(&goal_root));
        self.problem_def = Some(problem_def);
    }

    fn solve(&mut self, timeout: Duration) -> Result<Path<So3>, PlanningError> {
        let pd = self.problem_def.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let cones = self.cones.as_ref().ok_or(PlanningError::PlannerUninitialised)?.1;
        let ta = self.start_tree.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let tb = self.goal_tree.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let (start_time, mut rng) = (Instant::now(), rand::rng());
        let mut pa = vec![-1i64; self.max_nodes];
        let mut pb = vec![-1i64; self.max_nodes];
        let mut prm = ffi::oxg_plan_params {
            max_distance: self.max_distance,
            resolution: pd.space.get_longest_valid_segment_length() * 0.1,
            goal_radius: pd.goal.radius(),
            max_nodes: self.max_nodes as u64,
            stop_at_goal: 1,
            ..Default::default()
        };
        let target = so3_row(pd.goal.target());
        loop {
            let elapsed = start_time.elapsed();
            if elapsed > timeout {
                return Err(PlanningError::Timeout);
            }
            prm.timeout_s = (timeout - elapsed).as_secs_f64();
            let stream = so3_stream(pd, self.goal_bias, self.chunk, &mut rng);
            let mut r = ffi::oxg_plan_result::default();
            check(unsafe {
                ffi::oxg_so3_rrt_connect_solve(ta.0, tb.0, cones, &prm, target.as_ptr(),
                                               stream.as_ptr(), self.chunk as u64, pa.as_mut_ptr(),
                                               pb.as_mut_ptr(), &mut r)
            });
            if r.solved != 0 {
                // rrt_connect.rs:288-304: start-side chain, then the goal-side chain reversed
                let mut states = ta.chain(&pa, r.goal_node[0]);
                if r.goal_node[1] >= 0 {
                    let mut tail = tb.chain(&pb, r.goal_node[1]);
                    tail.reverse();
                    states.extend(tail.into_iter().skip(1));
                }
                return Ok(Path(states));
            }
            if r.status == 3 || r.status == 2 {
                return Err(PlanningError::Timeout);
            }
        }
    }
}

/// k nearest stored quaternions (`oxg_so3_knn`): rows of (index, distance), ascending.
pub fn so3_knn(store: ffi::oxg_store, queries: &[So3], k: usize) -> (Vec<u32>, Vec<f64>) {
    let rows: Vec<f64> = queries.iter().flat_map(so3_row).collect();
    let (mut idx, mut dist) = (vec![0u32; queries.len() * k], vec![0.0f64; queries.len() * k]);
    check(unsafe {
        ffi::oxg_so3_knn(store, rows.as_ptr(), queries.len() as u64, k as u32, std::ptr::null(),
                         idx.as_mut_ptr(), dist.as_mut_ptr())
    });
    (idx, dist)
}
