//! oxmpl-cuda — GPU backend behind oxmpl's existing traits.
//!
//! * [`GpuValidityChecker`] implements `oxmpl::base::validity::StateValidityChecker` for
//!   `RealVectorState` (declarative spheres / AABBs / capsules), so it drops into
//!   `Planner::setup(problem_def, Arc::new(checker))` of the stock planners unchanged.
//! * [`GpuRRT`] implements `oxmpl::base::planner::Planner` with the same constructor and
//!   `setup` / `solve(timeout)` contract as `oxmpl::geometric::RRT`, but keeps its tree in a
//!   GPU-resident store and runs the loop as one persistent kernel per chunk of pre-drawn
//!   samples (`oxg_rrt_solve`).  `GpuRRTConnect`, `GpuRRTStar` and `GpuPRM` follow the same
//!   pattern over `oxg_rrt_connect_solve`, `oxg_rrt_star_solve`, `oxg_prm_build_radius`.
//!
//! Sampling stays in Rust and uses exactly the reference's call sequence
//! (`rng.random_bool(goal_bias)` then `goal.sample_goal` or `space.sample_uniform`,
//! rrt.rs:177-184), which does not depend on the tree — so drawing a chunk up front changes
//! nothing the planner can observe.
pub mod ffi;

use std::{ffi::CStr, sync::Arc, time::{Duration, Instant}};

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
}
impl Drop for GpuValidityChecker {
    fn drop(&mut self) { unsafe { ffi::oxg_world_destroy(self.world); } }
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
    checker: Option<Arc<GpuValidityChecker>>,
    tree: Option<Store>,
}

impl<G: BallGoal> GpuRRT<G> {
    pub fn new(max_distance: f64, goal_bias: f64) -> Self {
        Self { max_distance, goal_bias, chunk: 1 << 14, max_nodes: 1 << 20, problem_def: None,
               checker: None, tree: None }
    }
    /// The GPU path needs the concrete checker (its device handle); `Planner::setup` below takes
    /// the trait object the reference's signature demands and this records the concrete one.
    pub fn set_gpu_checker(&mut self, checker: Arc<GpuValidityChecker>) { self.checker = Some(checker); }
}

impl<G: BallGoal> Planner<RealVectorState, RealVectorStateSpace, G> for GpuRRT<G> {
    fn setup(
        &mut self,
        problem_def: Arc<ProblemDefinition<RealVectorState, RealVectorStateSpace, G>>,
        _validity_checker: Arc<dyn StateValidityChecker<RealVectorState>>,
    ) {
        // rrt.rs:139-156: clear the tree, push the start state
        let tree = Store::new(problem_def.space.dimension);
        tree.push(&problem_def.start_states[0]);
        self.tree = Some(tree);
        self.problem_def = Some(problem_def);
    }

    fn solve(&mut self, timeout: Duration) -> Result<Path<RealVectorState>, PlanningError> {
        let pd = self.problem_def.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let vc = self.checker.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let tree = self.tree.as_ref().ok_or(PlanningError::PlannerUninitialised)?;
        let dim = pd.space.dimension;
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
            let mut stream = Vec::with_capacity(self.chunk * dim);
            for _ in 0..self.chunk {
                let q = if rng.random_bool(self.goal_bias) {
                    pd.goal.sample_goal(&mut rng).unwrap()
                } else {
                    pd.space.sample_uniform(&mut rng).unwrap()
                };
                stream.extend_from_slice(&q.values);
            }
            let mut r = ffi::oxg_plan_result::default();
            check(unsafe {
                ffi::oxg_rrt_solve(vc.handle(), tree.0, &prm, pd.goal.target().values.as_ptr(),
                                   stream.as_ptr(), self.chunk as u64, parents.as_mut_ptr(), &mut r)
            });
            if r.solved != 0 {
                // reconstruct_path (rrt.rs:118-128)
                let mut states = Vec::new();
                let mut i = r.goal_node[0];
                while i >= 0 {
                    states.push(tree.get(i as usize));
                    i = parents[i as usize];
                }
                states.reverse();
                return Ok(Path(states));
            }
            if r.status == 3 || r.status == 2 {
                return Err(PlanningError::Timeout);
            }
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
    fn drop(&mut self) { unsafe { ffi::oxg_so3_cones_destroy(self.cones); } }
}
impl StateValidityChecker<oxmpl::base::state::SO3State> for GpuConeChecker {
    fn is_valid(&self, state: &oxmpl::base::state::SO3State) -> bool {
        let v = [state.x, state.y, state.z, state.w];
        let mut out = 0u8;
        check(unsafe { ffi::oxg_so3_is_valid(self.cones, v.as_ptr(), 1, &mut out) });
        out != 0
    }
}
