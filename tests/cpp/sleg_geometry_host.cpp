// TEST INFRASTRUCTURE ONLY -- never linked into libamgb.so.
// Host-side geometry of the streaming legs (algebraic-multigrid_b200/csrc/stream_leg_api.hpp): how the
// lines of a level are cut into chunks, and which chunks are "edge chunks" of a sharded leg (the ones
// that wait for a neighbour and push to it).  tests/test_sleg_geometry_host.py checks the properties
// the fused halo push relies on by brute force.
#include "../../algebraic-multigrid_b200/csrc/stream_leg_api.hpp"

extern "C" {
// out: LJ, LJ_edge, n_chunks, n_warps, edge_lo_chunks, edge_hi_chunk0, expected_lo, expected_hi
void sleg_geometry(int n_lines, int m, int n_strips, int chunks, int edge_half, int NS, int lo_reach, int hi_reach,
                   int* out) {
  amgb::sleg::Params P{};
  P.n_lines = n_lines;
  P.m = m;
  P.n_strips = n_strips;
  P.set_chunks(chunks, edge_half != 0);
  amgb::sleg::Sync Y{};
  amgb::sleg::classify_edges(P, NS, lo_reach, hi_reach, Y);
  out[0] = P.LJ;
  out[1] = P.LJ_edge;
  out[2] = P.n_chunks;
  out[3] = P.n_warps;
  out[4] = Y.edge_lo_chunks;
  out[5] = Y.edge_hi_chunk0;
  out[6] = Y.expected[0];
  out[7] = Y.expected[1];
}
int sleg_chunk_begin(int n_lines, int n_strips, int chunks, int edge_half, int c) {
  amgb::sleg::Params P{};
  P.n_lines = n_lines;
  P.n_strips = n_strips;
  P.set_chunks(chunks, edge_half != 0);
  return P.chunk_begin(c);
}
int sleg_chunk_lines(int n_lines, int n_strips, int chunks, int edge_half, int c) {
  amgb::sleg::Params P{};
  P.n_lines = n_lines;
  P.n_strips = n_strips;
  P.set_chunks(chunks, edge_half != 0);
  return P.chunk_lines(c);
}
}
