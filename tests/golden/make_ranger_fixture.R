# Mint the one fixture this repository cannot produce itself (R / ranger are not in the build
# image): a ranger forest saved by R next to R's own predict() output, which would pin stage 2
# (forest mean and type-7 quantiles) against the reference implementation bit for bit.
#
#   Rscript tests/golden/make_ranger_fixture.R /path/to/RFSI/temp_croatia/temp_data
#
# Writes tests/golden/ranger_fixture.rda; tests/test_ranger_fixture.py picks it up when present
# (rfsi_b200/rda.py reads the file without R).
library(ranger); library(meteo); library(spacetime); library(sp)
args <- commandArgs(trailingOnly = TRUE)
load(file.path(args[1], "stfdf.rda"))
day <- stfdf[, 5]; day <- day[!is.na(day$TEMP), ]                       # 2_predictions.R:122-125
xy <- coordinates(day)
no <- near.obs(locations = day, observations = day, zcol = "TEMP", n.obs = 7)   # 1_models.R:1040-1045
X <- cbind(data.frame(x = xy[, 1], y = xy[, 2]), no)
df <- cbind(TEMP = day$TEMP, X)
set.seed(42)
model <- ranger(TEMP ~ ., data = df, num.trees = 250, mtry = 4, min.node.size = 6, quantreg = TRUE, seed = 42)
grid <- expand.grid(x = seq(min(xy[, 1]), max(xy[, 1]), length.out = 40), y = seq(min(xy[, 2]), max(xy[, 2]), length.out = 40))
gno <- near.obs(locations = grid, observations = as.data.frame(cbind(xy, TEMP = day$TEMP)), zcol = "TEMP", n.obs = 7)
G <- cbind(grid, gno)
pred <- predict(model, G)$predictions
q <- predict(model, G, type = "quantiles", quantiles = c(0.25, 0.5, 0.75))$predictions
tn <- predict(model, G, type = "terminalNodes")$predictions
stations <- as.data.frame(cbind(xy, TEMP = day$TEMP))
save(model, G, gno, pred, q, tn, stations, file = "tests/golden/ranger_fixture.rda", version = 2)
