# Parity harness for a machine that has R, meteo, nabor, ranger AND a B200 -- the one comparison
# this repository cannot run itself (no R in the build image).  It drives the reference's own call
# shapes through both implementations on the reference's Croatian station table:
#
#   Rscript rfsi_b200/R/tests/parity_against_meteo_ranger.R /path/to/RFSI/temp_croatia/temp_data
#
# and stops at the first difference beyond the tolerances of BASELINE.json's north_star:
# neighbour distances / observations 1e-12 relative (indices exact where distances are distinct),
# forest means and type-7 quantiles 1e-12 relative.
suppressMessages({library(meteo); library(ranger); library(sp); library(spacetime)})
args <- commandArgs(trailingOnly = TRUE)
dyn.load("rfsi_b200/R/src/rfsib200.so")
b200 <- new.env()
for (f in list.files("rfsi_b200/R/R", full.names = TRUE)) sys.source(f, envir = b200)

load(file.path(args[1], "stfdf.rda"))
rel <- function(a, b) max(abs(a - b) / pmax(1, abs(b)), na.rm = TRUE)
for (i in c(5, 13, 20, 28)) {                                     # the mapped days of 2_predictions.R:65
  st <- stfdf[, i]; st <- st[!is.na(st$TEMP), ]
  xy <- coordinates(st)
  grid <- expand.grid(x = seq(min(xy[, 1]), max(xy[, 1]), length.out = 200),
                      y = seq(min(xy[, 2]), max(xy[, 2]), length.out = 200))
  obs <- data.frame(x = xy[, 1], y = xy[, 2], TEMP = st$TEMP)
  ref <- meteo::near.obs(locations = grid, observations = obs, zcol = "TEMP", n.obs = 10)     # 2_predictions.R:160-165
  got <- b200$near.obs(locations = grid, observations = obs, zcol = "TEMP", n.obs = 10)
  stopifnot(identical(names(ref), names(got)), rel(as.matrix(got), as.matrix(ref)) <= 1e-12)
  # training shape: locations == observations, own observation excluded (1_models.R:1040-1045)
  ref1 <- meteo::near.obs(locations = obs, observations = obs, zcol = "TEMP", n.obs = 7)
  got1 <- b200$near.obs(locations = obs, observations = obs, zcol = "TEMP", n.obs = 7)
  stopifnot(rel(as.matrix(got1), as.matrix(ref1)) <= 1e-12)
  df <- cbind(obs, ref1)
  model <- ranger(TEMP ~ ., data = df, num.trees = 250, mtry = 4, min.node.size = 6, quantreg = TRUE, seed = 42)
  newdf <- cbind(grid, ref[, c(paste0("dist", 1:7), paste0("obs", 1:7))])
  fo <- b200$rfsi_import_ranger(model)
  p_ref <- predict(model, newdf)$predictions                                                    # 2_predictions.R:167
  q_ref <- predict(model, newdf, type = "quantiles", quantiles = c(0.25, 0.5, 0.75))$predictions  # 2_predictions.R:174-176
  p_got <- b200$predict_b200(fo, newdf)$predictions
  q_got <- b200$predict_b200(fo, newdf, type = "quantiles", quantiles = c(0.25, 0.5, 0.75))$predictions
  stopifnot(rel(p_got, p_ref) <= 1e-12, rel(q_got, q_ref) <= 1e-12)
  cat(sprintf("day %d: near.obs, predict and quantiles agree (%d stations, %d grid cells)\n", i, nrow(obs), nrow(grid)))
}
cat("parity with meteo / ranger: OK\n")
