// Distance from a packet to the next face of its (homologously expanding) propagation cell, for the three
// grid types, and the cell-change / escape bookkeeping.
// Follows reference grid.cc:2480-2755 (boundary_distance), 1412-1506 (expanding_shell_intersection),
// 1518-1525 (distance_cartesian_boundary), 1530-1553 (overshoot tolerance), 2461-2477 (snap_pos_to_cell),
// grid.h:114-137 (change_cell_or_escape), grid.cc:200-225 (cell coordinate helpers).
// Parity contract (BASELINE.json north_star): next-cell index bit-exact, distance <= 1e-12 relative.
#pragma once
#include "hd.h"
#include "options.h"
#include "packet.h"
#include "tables.h"
#include "vec.h"

namespace ab {

AHD int grid_ndim(const Tables& T) {
  return (T.grid_type == GRID_SPHERICAL1D) ? 1 : ((T.grid_type == GRID_CYLINDRICAL2D) ? 2 : 3);
}

AHD int coord_stride(const Tables& T, const int axis) {  // grid.cc:200-206
  int stride = 1;
  for (int a = 0; a < axis; ++a) {
    stride *= T.ncoord[a];
  }
  return stride;
}

AHD int cell_coordindex(const Tables& T, const int cellindex, const int axis) {  // grid.cc:209-211
  return (cellindex / coord_stride(T, axis)) % T.ncoord[axis];
}

AHD const double* coord_axis(const Tables& T, const int axis) {
  return (axis == 0) ? T.coord0 : ((axis == 1) ? T.coord1 : T.coord2);
}

AHD double cell_coordmin(const Tables& T, const int cellindex, const int axis) {  // grid.cc:215-217
  return coord_axis(T, axis)[cell_coordindex(T, cellindex, axis)];
}

AHD double cell_coordmax(const Tables& T, const int cellindex, const int axis) {  // grid.cc:221-225
  const int idx = cell_coordindex(T, cellindex, axis);
  return (idx < T.ncoord[axis] - 1) ? coord_axis(T, axis)[idx + 1] : T.rmax;
}

AHD double cellbound_tolerance(const double boundarypos) {  // grid.cc:1530-1532
  return dmax(10., fabs(boundarypos) * 1e-12);
}

template <bool UPPER>
AHD bool boundary_overshoot_within_tolerance(const Tables& T, const double pktposgridcoord,
                                             const double pktvelgridcoord, const double boundarypos_tmin,
                                             const double tstart) {  // grid.cc:1541-1553
  const double boundaryvel = boundarypos_tmin / T.tmin;
  const double boundarypos = boundaryvel * tstart;
  const double overshoot = UPPER ? (pktposgridcoord - boundarypos) : (boundarypos - pktposgridcoord);
  const bool movingtowards = UPPER ? (pktvelgridcoord > boundaryvel) : (pktvelgridcoord < boundaryvel);
  return movingtowards && (overshoot >= 0.) && (overshoot <= cellbound_tolerance(boundarypos));
}

AHD double distance_cartesian_boundary(const Tables& T, const double pktposgridcoord, const double pktvelgridcoord,
                                       const double cellboundarypos, const double tstart) {  // grid.cc:1518-1525
  return CLIGHT_PROP * (pktposgridcoord - (cellboundarypos / T.tmin * tstart)) /
         ((cellboundarypos / T.tmin) - pktvelgridcoord);
}

// forward distance to an expanding sphere (NDIM=3) or circle (NDIM=2); -1 if none (grid.cc:1412-1506)
template <bool UPPER, int NDIM>
AHD double expanding_shell_intersection(const double* pos, const double* dir, const double speed,
                                        const double shellradiuststart, const double tstart) {
  double dirdotdir = 0.;
  double dirdotpos = 0.;
  double posdotpos = 0.;
#pragma unroll
  for (int d = 0; d < NDIM; d++) {
    dirdotdir += dir[d] * dir[d];
    dirdotpos += dir[d] * pos[d];
    posdotpos += pos[d] * pos[d];
  }
  const double a = dirdotdir - pow2(shellradiuststart / tstart / speed);
  const double b = 2 * (dirdotpos - (pow2(shellradiuststart) / tstart / speed));
  const double c = posdotpos - pow2(shellradiuststart);
  const double discriminant = pow2(b) - (4 * a * c);

  if (discriminant < 0) {
    return -1;
  }
  if (discriminant > 0) {
    double dist1 = (-b + sqrt(discriminant)) / 2 / a;
    double dist2 = (-b - sqrt(discriminant)) / 2 / a;
    double posfinal1[NDIM];
    double posfinal2[NDIM];
    double dirdotpf1 = 0.;
    double dirdotpf2 = 0.;
    double len1sq = 0.;
    double len2sq = 0.;
#pragma unroll
    for (int d = 0; d < NDIM; d++) {
      posfinal1[d] = pos[d] + (dist1 * dir[d]);
      posfinal2[d] = pos[d] + (dist2 * dir[d]);
    }
#pragma unroll
    for (int d = 0; d < NDIM; d++) {
      dirdotpf1 += dir[d] * posfinal1[d];
      dirdotpf2 += dir[d] * posfinal2[d];
      len1sq += pow2(posfinal1[d]);
      len2sq += pow2(posfinal2[d]);
    }
    const double v_rad_shell = shellradiuststart / tstart;
    const double v_rad_final1 = dirdotpf1 * speed / sqrt(len1sq);
    const double v_rad_final2 = dirdotpf2 * speed / sqrt(len2sq);
    if constexpr (!UPPER) {
      if (v_rad_final1 > v_rad_shell) {
        dist1 = -1;
      }
      if (v_rad_final2 > v_rad_shell) {
        dist2 = -1;
      }
    } else {
      if (v_rad_final1 < v_rad_shell) {
        dist1 = -1;
      }
      if (v_rad_final2 < v_rad_shell) {
        dist2 = -1;
      }
    }
    if (dist1 < 0 && dist2 < 0) {
      return -1;
    }
    if (dist2 < 0) {
      return dist1;
    }
    if (dist1 < 0) {
      return dist2;
    }
    return dmin(dist1, dist2);
  }
  return -1.;  // tangential: ignored
}

struct BoundaryHit {
  double distance;
  int next_cellindex;
};

// innermost radius of a cell at tmin (grid.cc:228-253), for FORCE_SPHERICAL_ESCAPE_SURFACE
AHD double cell_r_inner(const Tables& T, const int cellindex) {
  if (T.grid_type == GRID_SPHERICAL1D) {
    return cell_coordmin(T, cellindex, 0);
  }
  auto axis_mindist = [](const double cmin, const double cmax) {
    return (cmin <= 0. && cmax >= 0.) ? 0. : dmin(fabs(cmin), fabs(cmax));
  };
  if (T.grid_type == GRID_CYLINDRICAL2D) {
    const double rcyl_inner = cell_coordmin(T, cellindex, 0);
    const double z_inner = axis_mindist(cell_coordmin(T, cellindex, 1), cell_coordmax(T, cellindex, 1));
    return sqrt(pow2(rcyl_inner) + pow2(z_inner));
  }
  const double x_inner = axis_mindist(cell_coordmin(T, cellindex, 0), cell_coordmax(T, cellindex, 0));
  const double y_inner = axis_mindist(cell_coordmin(T, cellindex, 1), cell_coordmax(T, cellindex, 1));
  const double z_inner = axis_mindist(cell_coordmin(T, cellindex, 2), cell_coordmax(T, cellindex, 2));
  return sqrt(pow2(x_inner) + pow2(y_inner) + pow2(z_inner));
}

AHD BoundaryHit boundary_distance(const Tables& T, const double* dir, const double* pos, const double tstart,
                                  const int cellindex) {
  if constexpr (opt::FORCE_SPHERICAL_ESCAPE_SURFACE) {
    if (cell_r_inner(T, cellindex) > T.rmax) {
      return {0., -99};
    }
  }
  const int gridtype = static_cast<int>(T.grid_type);

  double distance = DBL_MAX_;
  int next_cellindex = -1;

  if (gridtype == GRID_CARTESIAN3D) {
    // grid.cc:2698-2735. Strides: 1, nx, nx*ny
    int stride = 1;
    int rem = cellindex;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const int n = T.ncoord[d];
      const int idx = rem % n;
      rem /= n;
      const double* coords = coord_axis(T, d);
      const double cmin = coords[idx];
      const double cmax = (idx < n - 1) ? coords[idx + 1] : T.rmax;
      const double posd = pos[d];
      const double veld = dir[d] * CLIGHT_PROP;
      if (veld > (cmax / T.tmin)) {
        const double dmaxb = boundary_overshoot_within_tolerance<true>(T, posd, veld, cmax, tstart)
                                 ? 0.
                                 : distance_cartesian_boundary(T, posd, veld, cmax, tstart);
        if ((dmaxb >= 0.) && (dmaxb < distance)) {
          distance = dmaxb;
          next_cellindex = (idx == (n - 1)) ? -99 : cellindex + stride;
        }
      } else if (veld < (cmin / T.tmin)) {
        const double dminb = boundary_overshoot_within_tolerance<false>(T, posd, veld, cmin, tstart)
                                 ? 0.
                                 : distance_cartesian_boundary(T, posd, veld, cmin, tstart);
        if ((dminb >= 0.) && (dminb < distance)) {
          distance = dminb;
          next_cellindex = (idx == 0) ? -99 : cellindex - stride;
        }
      }
      stride *= n;
    }
  } else if (gridtype == GRID_SPHERICAL1D) {
    // grid.cc:2571-2601
    const double posr = vec_len3(pos);                                 // get_gridcoords_from_xyz
    const double velr = dot3(pos, dir) / posr * CLIGHT_PROP;           // get_gridcoords_vel_from_xyz_pos_dir
    const int idx = cell_coordindex(T, cellindex, 0);
    const double cmin = T.coord0[idx];
    const double cmax = (idx < T.ncoord[0] - 1) ? T.coord0[idx + 1] : T.rmax;
    const double speed = vec_len3(dir) * CLIGHT_PROP;

    const double r_outer = cmax * tstart / T.tmin;
    const double d_up = boundary_overshoot_within_tolerance<true>(T, posr, velr, cmax, tstart)
                            ? 0.
                            : expanding_shell_intersection<true, 3>(pos, dir, speed, r_outer, tstart);
    if ((d_up >= 0.) && (d_up < distance)) {
      distance = d_up;
      next_cellindex = (idx == (T.ncoord[0] - 1)) ? -99 : cellindex + 1;
    }
    const double r_inner = cmin * tstart / T.tmin;
    if (r_inner > 0.) {
      const double d_lo = boundary_overshoot_within_tolerance<false>(T, posr, velr, cmin, tstart)
                              ? 0.
                              : expanding_shell_intersection<false, 3>(pos, dir, speed, r_inner, tstart);
      if ((d_lo >= 0.) && (d_lo < distance)) {
        distance = d_lo;
        next_cellindex = (idx == 0) ? -99 : cellindex - 1;
      }
    }
  } else {
    // CYLINDRICAL2D: grid.cc:2602-2696. coordinate 0 = cylindrical radius, coordinate 1 = z
    const double posrcyl = sqrt(pow2(pos[0]) + pow2(pos[1]));
    const double posz = pos[2];
    const double velrcyl = ((pos[0] * dir[0]) + (pos[1] * dir[1])) / posrcyl * CLIGHT_PROP;
    const double velz = dir[2] * CLIGHT_PROP;
    const int nr = T.ncoord[0];
    const int nz = T.ncoord[1];
    const int ir = cellindex % nr;
    const int iz = (cellindex / nr) % nz;
    const double rmin = T.coord0[ir];
    const double rmaxc = (ir < nr - 1) ? T.coord0[ir + 1] : T.rmax;
    const double zmin = T.coord1[iz];
    const double zmax = (iz < nz - 1) ? T.coord1[iz + 1] : T.rmax;

    const double posnoz[2] = {pos[0], pos[1]};
    const double dirxylen = sqrt(pow2(dir[0]) + pow2(dir[1]));
    const double xyspeed = dirxylen * CLIGHT_PROP;

    if (dirxylen > 0.) {
      const double dirnoz[2] = {dir[0] / dirxylen, dir[1] / dirxylen};
      const double r_outer = rmaxc * tstart / T.tmin;
      const double d_rcyl_up = boundary_overshoot_within_tolerance<true>(T, posrcyl, velrcyl, rmaxc, tstart)
                                   ? 0.
                                   : expanding_shell_intersection<true, 2>(posnoz, dirnoz, xyspeed, r_outer, tstart);
      if (d_rcyl_up >= 0.) {
        const double d_z = d_rcyl_up / xyspeed * dir[2] * CLIGHT_PROP;
        const double d_tot = sqrt(pow2(d_rcyl_up) + pow2(d_z));
        if ((d_tot >= 0.) && (d_tot < distance)) {
          distance = d_tot;
          next_cellindex = (ir == (nr - 1)) ? -99 : cellindex + 1;
        }
      }
      const double r_inner = rmin * tstart / T.tmin;
      if (r_inner > 0) {
        const double d_rcyl_lo = boundary_overshoot_within_tolerance<false>(T, posrcyl, velrcyl, rmin, tstart)
                                     ? 0.
                                     : expanding_shell_intersection<false, 2>(posnoz, dirnoz, xyspeed, r_inner, tstart);
        if (d_rcyl_lo >= 0.) {
          const double d_z = d_rcyl_lo / xyspeed * dir[2] * CLIGHT_PROP;
          const double d_tot = sqrt(pow2(d_rcyl_lo) + pow2(d_z));
          if ((d_tot >= 0.) && (d_tot < distance)) {
            distance = d_tot;
            next_cellindex = (ir == 0) ? -99 : cellindex - 1;
          }
        }
      }
    } else {
      // moving exactly along z (grid.cc:2652-2670): only the expanding inner boundary can catch up
      if (rmin > 0.) {
        const double d_lo = boundary_overshoot_within_tolerance<false>(T, posrcyl, velrcyl, rmin, tstart)
                                ? 0.
                                : ((posrcyl * T.tmin / rmin) - tstart) * CLIGHT_PROP;
        if ((d_lo >= 0.) && (d_lo < distance)) {
          distance = d_lo;
          next_cellindex = (ir == 0) ? -99 : cellindex - 1;
        }
      }
    }

    // z boundaries are Cartesian (grid.cc:2672-2696)
    if (velz > (zmax / T.tmin)) {
      const double d_up = boundary_overshoot_within_tolerance<true>(T, posz, velz, zmax, tstart)
                              ? 0.
                              : distance_cartesian_boundary(T, posz, velz, zmax, tstart);
      if ((d_up >= 0.) && (d_up < distance)) {
        distance = d_up;
        next_cellindex = (iz == (nz - 1)) ? -99 : cellindex + nr;
      }
    } else if (velz < (zmin / T.tmin)) {
      const double d_lo = boundary_overshoot_within_tolerance<false>(T, posz, velz, zmin, tstart)
                              ? 0.
                              : distance_cartesian_boundary(T, posz, velz, zmin, tstart);
      if ((d_lo >= 0.) && (d_lo < distance)) {
        distance = d_lo;
        next_cellindex = (iz == 0) ? -99 : cellindex - nr;
      }
    }
  }

  if (distance > T.max_path_step) {  // grid.cc:2750-2752
    return {T.max_path_step, cellindex};
  }
  return {distance, next_cellindex};
}

// clamp the position into the new cell after a crossing; Cartesian grids only (grid.cc:2461-2477)
AHD void snap_pos_to_cell(const Tables& T, double* pos, const double time, const int cellindex) {
  if (T.grid_type != GRID_CARTESIAN3D) {
    return;
  }
  int rem = cellindex;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int n = T.ncoord[d];
    const int idx = rem % n;
    rem /= n;
    const double* coords = coord_axis(T, d);
    const double cellposmin = coords[idx] / T.tmin * time;
    const double cellposmax = (idx < (n - 1)) ? coords[idx + 1] / T.tmin * time : T.rmax / T.tmin * time;
    const double x = pos[d];
    pos[d] = (x < cellposmin) ? cellposmin : ((cellposmax < x) ? cellposmax : x);  // std::clamp
  }
}

// grid.h:114-137
AHD void change_cell_or_escape(Pkt& p, const Ctx& c, const int next_cellindex) {
  if (next_cellindex >= 0) {
    if (next_cellindex != p.cellindex) {
      snap_pos_to_cell(c.T, p.pos, p.prop_time, next_cellindex);
    }
    p.cellindex = next_cellindex;
    c.count<CNT_CELLCROSSINGS>();
  } else {
    c.T.pkt.escape_type[c.ip] = p.type;
    c.T.pkt.escape_time[c.ip] = static_cast<float>(p.prop_time);
    p.type = TYPE_ESCAPE;
    c.count<CNT_PKTESCAPES>();
  }
}

}  // namespace ab
