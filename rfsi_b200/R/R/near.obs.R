# near.obs -- drop-in for meteo::near.obs as the reference scripts call it
# (simulation/simulation.R:191-196; sim_500.R:134-139; temp_croatia/2_predictions.R:160-165;
#  prcp_catalonia/5_prcp_prediction.R:278-283; temp_croatia/1_models.R:1040-1045).
# Same argument names, same defaults, same return value: a data.frame with columns
# dist1..distN then obs1..obsN.  The search runs on the GPU through .Call(C_rfsi_near_obs).
near.obs <- function(locations, locations.x.y = c(1, 2), observations, observations.x.y = c(1, 2),
                     zcol = 3, n.obs = 10, rm.dupl = TRUE, avg = FALSE, increment, range,
                     direct = FALSE, idw = FALSE, idw.p = 2, device = 0L) {
  if (avg || direct || idw) stop("near.obs (rfsi_b200): avg/direct/idw are not supported")
  xy <- function(obj, cols) {
    if (inherits(obj, "Spatial")) return(unname(as.matrix(sp::coordinates(obj))[, 1:2, drop = FALSE]))
    unname(as.matrix(as.data.frame(obj)[, cols, drop = FALSE]))
  }
  loc <- xy(locations, locations.x.y)
  obs <- xy(observations, observations.x.y)
  z <- if (inherits(observations, "Spatial")) observations@data[[zcol]] else as.data.frame(observations)[[zcol]]
  storage.mode(loc) <- "double"; storage.mode(obs) <- "double"
  # the engine fills the 2 * n.obs columns of the returned data.frame in place (no matrix, no as.data.frame copy)
  .Call(C_rfsi_near_obs_df, loc, obs, as.double(z), as.integer(n.obs), as.logical(rm.dupl), as.integer(device))
}

# near.obs.days -- all days of a station table in one engine call: what
#   foreach(t = time) %dopar% near.obs(locations = day_df, observations = day_df, zcol, n.obs)
# builds day by day (temp_croatia/1_models.R:1034-1047).  z: stations x days matrix (NA = no
# observation that day).  locations = NULL: the stations themselves (own observation excluded).
# Returns list(dist, obs): arrays [location, neighbour, day].
near.obs.days <- function(observations.xy, z, n.obs = 10, locations.xy = NULL, device = 0L) {
  oxy <- unname(as.matrix(observations.xy)); storage.mode(oxy) <- "double"
  z <- unname(as.matrix(z)); storage.mode(z) <- "double"
  lxy <- if (is.null(locations.xy)) NULL else { m <- unname(as.matrix(locations.xy)); storage.mode(m) <- "double"; m }
  res <- .Call(C_rfsi_near_obs_days, lxy, oxy, z, as.integer(n.obs), as.integer(device))
  npix <- if (is.null(lxy)) nrow(oxy) else nrow(lxy)
  list(dist = array(res[[1]], c(npix, n.obs, ncol(z))), obs = array(res[[2]], c(npix, n.obs, ncol(z))))
}
